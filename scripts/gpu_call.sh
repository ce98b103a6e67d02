(timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02m_tests.log; tail -8 gpurun_out/r02m_tests.log
