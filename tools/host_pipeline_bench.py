"""Host input-pipeline throughput: PNG decode + raw label read of the eval loader (kfnet_b200/eval.py:
load_image / load_label, replacing train.py:195-239's 1-thread TF queue) over a thread pool.  CPU only.
  python tools/host_pipeline_bench.py [--frames 64] [--threads 1,2,4,8,16]"""
import argparse
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kfnet_b200 import eval as E, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--threads", default="1,2,4,8,16")
    a = ap.parse_args()
    import cv2
    d = tempfile.mkdtemp(prefix="kfn_host_")
    seq = synth.make_sequence(8, seed=3)
    imgs, labs = [], []
    for i in range(a.frames):
        p = os.path.join(d, "f%04d.png" % i)
        cv2.imwrite(p, np.roll(seq[i % 8], i, axis=1))
        q = os.path.join(d, "l%04d.bin" % i)
        np.random.default_rng(i).normal(size=(480, 640, 4)).astype("<f4").tofile(q)
        imgs.append(p); labs.append(q)
    print("PNG %.0f KB/frame, label %.1f MB/frame, %d logical cores" % (os.path.getsize(imgs[0]) / 1e3, os.path.getsize(labs[0]) / 1e6, os.cpu_count()))
    for nt in [int(x) for x in a.threads.split(",")]:
        with ThreadPoolExecutor(max_workers=nt) as pool:
            for what, fn, paths in (("decode", E.load_image, imgs), ("decode+label", None, None)):
                t0 = time.perf_counter()
                if fn:
                    list(pool.map(fn, paths))
                else:
                    fi = [pool.submit(E.load_image, p) for p in imgs]
                    fl = [pool.submit(E.load_label, p, (480, 640)) for p in labs]
                    [f.result() for f in fi + fl]
                dt = time.perf_counter() - t0
                print("threads %2d  %-13s %7.1f frames/s" % (nt, what, a.frames / dt))


if __name__ == "__main__":
    main()
