timeout 120 python tools/prof_decode.py --presets 0,7 --seconds 10 --reps 5 2>&1 | tail -3
timeout 200 python tools/prof_decode.py --presets 7 --seconds 60 --channels 8 --bits 24 --rate 96000 --reps 3 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tput.py -m gpu -x -q 2>&1 | tail -3
