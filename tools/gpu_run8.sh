mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/r2i_tests.log
