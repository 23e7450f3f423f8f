python -m pytest tests -m gpu -q -k "scomb or NOCOMB or perline or kit or image or component" > gpurun_out/r2_tests2.log 2>&1; tail -40 gpurun_out/r2_tests2.log
