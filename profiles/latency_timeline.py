import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spl_slam_b200 as S
import bench
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (752, 480)
img = bench.synth_image(W, H, 0)
cl = S.Context(0, priority=1)
le = S.Lineextractor(200, 2, 0, 1.1, 0.8, 2.2, 12.5, 1.0, 0.8, 1024, 0.0, ctx=cl)
for _ in range(5): le.ComputeLsdWithLbd(img)
cl.profile_enable(True)
cl.timer_start()
le.ComputeLsdWithLbd(img)
tl = cl.profile_timeline(cl)
for n, a, b in tl: print("%-28s %7.3f %7.3f  (%.3f)" % (n, a, b, b - a))
