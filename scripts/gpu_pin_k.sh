for k in 64 128 256; do
  MIDAS_B200_NBR_K=$k timeout 600 python scripts/pin_ab.py 2>/dev/null
done
