for f in 32 128 512; do echo "FIBERS=$f"; SMALT_B200_FIBERS=$f SMALT_B200_BLOCK=8192 python tools/paired_check.py 20000 4 4 2>&1 | grep -m2 "pairs\|fiber"; done
