#!/bin/bash
# Round 2, tenth GPU visit (1 GPU): the path tracer — parity tests of the new kernels, the C++ adapter, a first throughput line.
tag=${1:-r2j}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_pt.py tests/test_host_cpp.py tests/test_film_io.py -m gpu -q -s ) > gpurun_out/${tag}_pt_tests.log 2>&1
tail -6 gpurun_out/${tag}_pt_tests.log
grep -h "pt product parity\|FAILED\|Error\|assert" gpurun_out/${tag}_pt_tests.log | cut -c1-300 | head -40
timeout 300 python tools/pt_throughput.py C3 8 > gpurun_out/${tag}_pt_c3.json 2> gpurun_out/${tag}_pt_c3.err; tail -3 gpurun_out/${tag}_pt_c3.json; tail -3 gpurun_out/${tag}_pt_c3.err
exit 0
