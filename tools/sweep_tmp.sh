for lib in _s16 _t256; do for mc in 128 256 512 1024; do for div in 74 148 296 1024; do for mx in 4 8 16; do
echo -n "lib=$lib min=$mc div=$div max=$mx: "; DRAW_B200_SPLIT_MIN_COST=$mc DRAW_B200_SPLIT_DIV=$div DRAW_B200_SPLIT_MAX=$mx AB_STEPS=100 tools/ab_quick.sh c3 -- "$lib" | grep fps | python -c "
import sys,re
l=sys.stdin.read(); m=re.search(r'us ([\d.]+) .*k_tile.: ([\d.]+)',l); print(m.group(1), m.group(2))"
done; done; done; done
