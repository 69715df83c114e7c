"""bench.py generates its synthetic inputs with numpy / scipy only (nothing under oracle/ may run outside the cpu_baseline /
reference legs); the generator must produce exactly the images the parity tests use."""
import numpy as np


def test_bench_generator_equals_oracle_generator(oracle):
    import bench
    for (w, h, seed) in [(752, 480, 0), (640, 480, 7), (320, 240, 3)]:
        assert np.array_equal(bench.synth_image(w, h, seed), oracle.synth_image(w, h, seed))
    assert bench._gauss_kernel_q8(7, 2.0) == [18, 34, 48, 56, 48, 34, 18]
    assert bench._gauss_kernel_q8(5, 1.0) == [14, 62, 104, 62, 14]


def test_bench_has_no_oracle_import_outside_baseline_legs():
    import os, re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    # the only place that touches oracle/: cpu_reference_throughput (the cpu_baseline / --impl reference legs)
    uses = [m.start() for m in re.finditer(r"from oracle import", src)]
    body = src[src.index("def cpu_reference_throughput"):src.index("def cpu_knn2")]
    assert len(uses) == 1 and "from oracle import" in body
