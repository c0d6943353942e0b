timeout 200 python tools/prof_decode.py --presets 7 --seconds 60 --channels 8 --bits 24 --rate 96000 --reps 3 2>&1 | tail -2
timeout 120 python tools/prof_decode.py --presets 0,7 --seconds 10 --reps 5 2>&1 | tail -2
