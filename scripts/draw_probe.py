"""Where does the bit-parity (host) subset draw spend its time?  Run on the GPU box (real cores):
    python scripts/draw_probe.py
Prints the serial stream walk (cg_host_legacy_skip), the threaded draw at several worker counts, and a breakdown of
GraspPredicter.predict_batch(subsample="host") into producer (draw) time and total wall clock."""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from catgrasp_b200.predicter import GraspPredicter, _LegacyDraw
    out = {"cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        out["model"] = [l.split(":")[1].strip() for l in txt.splitlines() if l.startswith("model name")][0]
        out["flags_avx2"] = " avx2 " in txt
        out["flags_avx512f"] = " avx512f " in txt
    except Exception as e:
        out["cpuinfo_error"] = str(e)
    np.random.seed(0)
    out["skip"], out["draw"] = [], []
    for M, count in ((20000, 1024), (3000, 2048)):
        d = _LegacyDraw()
        d.skip(M, 1024, 16)
        t = time.perf_counter()
        d.skip(M, 1024, count)
        dt = time.perf_counter() - t
        out["skip"].append({"M": M, "count": count, "ms": 1e3 * dt, "ns_per_elem": 1e9 * dt / count / M})
        for nt in (1, 2, 4, 8, 16, 32, 64):
            d = _LegacyDraw()
            buf = np.empty((count, 1024), np.int32)
            d.draw(M, 1024, 64, out=buf[:64], nthreads=nt)
            t = time.perf_counter()
            c0 = time.process_time()
            d.draw(M, 1024, count, out=buf, nthreads=nt)
            dt = time.perf_counter() - t
            out["draw"].append({"M": M, "count": count, "nthreads": nt, "ms": 1e3 * dt, "cpu_ms": 1e3 * (time.process_time() - c0),
                                "us_per_candidate": 1e6 * dt / count})
    import torch
    if torch.cuda.is_available():
        from catgrasp_b200.synthetic import make_candidates, make_pile, write_artifacts
        acc = {"draw_s": 0.0, "calls": 0}
        orig = _LegacyDraw.draw

        def timed(self, *a, **k):
            t = time.perf_counter()
            r = orig(self, *a, **k)
            acc["draw_s"] += time.perf_counter() - t
            acc["calls"] += 1
            return r
        _LegacyDraw.draw = timed
        out["predict_batch"] = []
        with tempfile.TemporaryDirectory() as td:
            adir = write_artifacts(os.path.join(td, "artifacts-47"), "cls", n_pts=1024, seed=0)
            with contextlib.redirect_stdout(io.StringIO()):
                gp = GraspPredicter("nut", artifact_dir=adir, device=0)
            for M, B in ((3000, 1024), (20000, 4096)):
                scene = make_pile(M, n_objects=4 if M < 10000 else 12, seed=5)
                data = {"cloud_xyz": scene["cloud_xyz"], "cloud_normal": scene["cloud_normal"]}
                poses = list(make_candidates(scene["cloud_xyz"], scene["cloud_normal"], B, seed=6))
                for chunk in (256, 512, 1024, 4096):
                    gp.chunk = chunk
                    gp.predict_batch(data, poses, subsample="host")
                    acc.update(draw_s=0.0, calls=0)
                    t = time.perf_counter()
                    gp.predict_batch(data, poses, subsample="host")
                    dt = time.perf_counter() - t
                    out["predict_batch"].append({"M": M, "B": B, "chunk": chunk, "ms": 1e3 * dt, "draw_ms": 1e3 * acc["draw_s"],
                                                 "draw_calls": acc["calls"], "cand_per_s": B / dt})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
