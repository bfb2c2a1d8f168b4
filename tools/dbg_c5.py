import sys, numpy as np
sys.path.insert(0,'/root/repo')
import draw_b200
from draw_b200 import synthetic
for (nt,nph,W,H) in ((250,200,1920,1080),(64,48,1280,720)):
    t = synthetic.torus(nt, nph, texture=synthetic.checker_material())
    s, c = draw_b200.Scene(W,H), draw_b200.Canvas(W,H); c.init_depth(100000.0); s.add_obj(t)
    for k in range(3):
        s.render(c); d = c.depth()
        print(nt, nph, 'frame', k, c.last_frame_stats(), 'covered', int((d<100000).sum()))
