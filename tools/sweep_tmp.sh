for sets in 2 3 4 6 8; do echo -n "sets=$sets: "; DRAW_B200_SETS=$sets tools/ab_quick.sh c3 -- "" | grep fps | cut -c1-48; done
for n in 148 296 1184; do echo -n "clear_ctas=$n: "; DRAW_B200_CLEAR_CTAS=$n tools/ab_quick.sh c3 -- "" | grep fps | cut -c1-48; done
for p in 1; do echo -n "prio=$p: "; DRAW_B200_PRIO=$p tools/ab_quick.sh c3 -- "" | grep fps | cut -c1-48; done
for m in 64 512 2048; do echo -n "split_min_cost=$m: "; DRAW_B200_SPLIT_MIN_COST=$m tools/ab_quick.sh c3 -- "" | grep fps | cut -c1-48; done
for m in 148 1024; do echo -n "split_div=$m: "; DRAW_B200_SPLIT_DIV=$m tools/ab_quick.sh c3 -- "" | grep fps | cut -c1-48; done
