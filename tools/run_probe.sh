for t in 150 250 350 400 600; do
echo "== GEGP_BIG_TILES=$t"
GEGP_BIG_TILES=$t python tools/one_eval.py 500 10 1 5 | tail -2
done
