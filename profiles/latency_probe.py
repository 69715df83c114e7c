"""Single-frame latency of the drop-in calls (what Frame::Frame does: ORB thread + line thread, src/Frame.cc:301-304)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
import spl_slam_b200 as S
from oracle import oracle as O

for (w, h, nf, nl, sig, dens) in [(640, 480, 1000, 600, 0.6, 0.6), (752, 480, 1200, 200, 0.8, 0.8), (1241, 376, 2000, 800, 0.6, 0.6)]:
    img = O.synth_image(w, h, 0)
    co, cl = S.Context(0), S.Context(0, priority=1)
    orb = S.ORBextractor(nf, 1.2, 8, 20, 7, ctx=co)
    le = S.Lineextractor(nl, 2, 0, 1.1, sig, 2.2, 12.5, 1.0, dens, 1024, 0.0, ctx=cl)
    pool = ThreadPoolExecutor(2)
    def both():
        f = pool.submit(orb, img)
        r = le.ComputeLsdWithLbd(img)
        return f.result(), r
    for _ in range(5): both()
    ts, to, tl = [], [], []
    for _ in range(30):
        t0 = time.perf_counter(); both(); ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); orb(img); to.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); le.ComputeLsdWithLbd(img); tl.append(time.perf_counter() - t0)
    # CPU oracle, single thread each, two threads like the reference
    ox = O.ORBextractor(nf, 1.2, 8, 20, 7); prm = O.line_params(nl, 2, 0, 1.1, sig, 2.2, 12.5, 1.0, dens, 1024, 0.0)
    def cpu_both():
        f = pool.submit(ox, img)
        r = O.line_extract(prm, img)
        return f.result(), r
    cpu_both()
    tc = []
    for _ in range(5):
        t0 = time.perf_counter(); cpu_both(); tc.append(time.perf_counter() - t0)
    print("%dx%d: GPU frame (ORB || lines) %.2f ms (ORB alone %.2f, lines alone %.2f); CPU oracle, two threads %.1f ms" %
          (w, h, 1e3 * np.median(ts), 1e3 * np.median(to), 1e3 * np.median(tl), 1e3 * np.median(tc)), flush=True)
