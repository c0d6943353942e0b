timeout 120 python tools/prof_decode.py --presets 0,4,7 --seconds 10 --reps 5 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tput.py tests/test_api_behaviour.py -m gpu -x -q 2>&1 | tail -5
