export LINNE_B200_PIPELINE=1
timeout 200 python tools/prof_tput.py --preset 7 --blocks 15488 --reps 1 2>&1 | tail -3
