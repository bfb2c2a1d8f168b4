python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for n in 296 222 148; do for sets in 4 6; do echo "tile_ctas=$n sets=$sets"; DRAW_B200_TILE_CTAS=$n DRAW_B200_SETS=$sets tools/ab_quick.sh c3 c2 c4 -- "" | grep fps | cut -c1-70; done; done
