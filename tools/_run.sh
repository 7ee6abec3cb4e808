timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/practical_bar.py > gpurun_out/practical_bar.json 2> gpurun_out/practical_bar.err; cat gpurun_out/practical_bar.json
