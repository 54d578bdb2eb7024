#!/usr/bin/env python3
"""Generate the golden frames under tests/golden/ from the UNMODIFIED reference (oracle/_ref builds).

Run where /root/reference is mounted:   python tests/golden/make_golden.py
Every case is one run of `renderer_<variant> -b -n N -m <mode> [-w] <model>` (the reference's own
benchmark orbit); the presented frames are stored losslessly (npz, zlib) with their SHA-256.
The reference itself ships no golden vectors or tests, so these are the pin.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyport  # noqa: E402

W, H = 320, 240

# name: (model, mode, frames, two_lights, variant kwargs, env)
CASES = {
    # ray tracer (reference defaults: shadows + reflections + phong normal)
    "rt_torus_m9": ("torus.ply", 9, [0, 37], False, {}, {}),
    "rt_chess_m9": ("chessboard.tri", 9, [0, 99], False, {}, {}),
    "rt_dragon_m9": ("dragon_vis.ply", 9, [1], False, {}, {}),
    "rt_train_m9_w": ("trainColor.tri", 9, [5], True, {}, {}),
    "rt_torus_m0": ("torus.ply", 0, [3], False, {}, {}),
    "rt_chess_m9_norefl": ("chessboard.tri", 9, [0, 12], False, {"no_reflections": True}, {}),
    "rt_torus_m9_ao16": ("torus.ply", 9, [2], False, {"ao": 16}, {"OMP_NUM_THREADS": "4"}),
    "rt_chess_m9_ao16": ("chessboard.tri", 9, [2], False, {"ao": 16}, {"OMP_NUM_THREADS": "4"}),
    # rasteriser: strict build, single-threaded so the (benign) Z-buffer races of the reference cannot occur
    "ras_torus_m1": ("torus.ply", 1, [0, 37], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_torus_m2": ("torus.ply", 2, [0, 37], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_torus_m3": ("torus.ply", 3, [0, 37], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_statue_m4": ("statue.ply", 4, [0, 50], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_statue_m5": ("statue.ply", 5, [0, 50], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_statue_m6": ("statue.ply", 6, [0, 50], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_statue_m7": ("statue.ply", 7, [0, 50], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_statue_m8": ("statue.ply", 8, [0, 50], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_chess_m6": ("chessboard.tri", 6, [0], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_chess_m8_w": ("chessboard.tri", 8, [10], True, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_train_m8": ("trainColor.tri", 8, [0, 20], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_chess_m3": ("chessboard.tri", 3, [0], False, {}, {"OMP_NUM_THREADS": "1"}),
    "ras_dragon_m2": ("dragon_vis.ply", 2, [4], False, {}, {"OMP_NUM_THREADS": "1"}),
    # MLAA post filter (built --enable-mlaa equivalent)
    "mlaa_statue_m6": ("statue.ply", 6, [0], False, {"mlaa": True}, {"OMP_NUM_THREADS": "1"}),
    "mlaa_chess_m9": ("chessboard.tri", 9, [0], False, {"mlaa": True}, {}),
    "mlaa_train_m8": ("trainColor.tri", 8, [7], False, {"mlaa": True}, {"OMP_NUM_THREADS": "1"}),
}


# BASELINE.json's configurations at their FULL sizes (frames {0, 1, 37, 99} of the -b orbit). A 4K frame is
# 33 MB, so only the SHA-256 (and the count of non-black pixels) is stored: tests compare digests; on a mismatch the CPU
# restatement (pinned to the same digests) supplies the pixels to diff against.
# name: (model, mode, width, height, frames, variant kwargs, env)
FULL_FRAMES = [0, 1, 37, 99]
FULL_CASES = {
    "full_c2_chess_m9_norefl": ("chessboard.tri", 9, 1920, 1080, FULL_FRAMES, {"no_reflections": True}, {}),
    "full_c3_dragon_m9_ao16": ("dragon_vis.ply", 9, 1920, 1080, FULL_FRAMES, {"ao": 16}, {"OMP_NUM_THREADS": "8"}),
    "full_c4_statue_m5_mlaa": ("statue.ply", 5, 3840, 2160, FULL_FRAMES, {"mlaa": True}, {"OMP_NUM_THREADS": "1"}),
    "full_c4_statue_m6_mlaa": ("statue.ply", 6, 3840, 2160, FULL_FRAMES, {"mlaa": True}, {"OMP_NUM_THREADS": "1"}),
    "full_c5_chess_m9_ao16": ("chessboard.tri", 9, 3840, 2160, FULL_FRAMES, {"ao": 16}, {"OMP_NUM_THREADS": "8"}),
}


def full_size(only, index):
    import time
    for name, (model, mode, w, h, frames, kw, env) in FULL_CASES.items():
        if only and name not in only:
            continue
        if not pyport.have_ref(w, h, **kw):
            pyport.build_ref(w, h, **kw)
        t0 = time.time()
        imgs, _ = pyport.run_ref(pyport.model_path(model), mode, w, h, frames, env=env, **kw)
        index[name] = {"model": model, "mode": mode, "frames": frames, "two_lights": False, "variant": kw,
                       "width": w, "height": h, "digest_only": True,
                       "lit_pixels": {str(k): int((v != 0).sum()) for k, v in imgs.items()},
                       "sha256": {str(k): hashlib.sha256(v.tobytes()).hexdigest() for k, v in imgs.items()}}
        print(name, index[name]["lit_pixels"], f"{time.time() - t0:.1f} s")


# The AO caveat of SURVEY.md section 8c: the parity oracle draws AO samples from a per-pixel seeded stream (the one permitted
# delta). These cases keep the reference's libc rand() (single thread, so the run is repeatable) - NOT parity targets, only the
# other side of a statistical comparison: same image up to sampling noise, no bias.
STAT_CASES = {
    "stat_chess_m9_ao16_libcrand": ("chessboard.tri", 9, [2], {"ao": 16, "libc_rand": True}, {"OMP_NUM_THREADS": "1"}, "rt_chess_m9_ao16"),
    "stat_torus_m9_ao16_libcrand": ("torus.ply", 9, [2], {"ao": 16, "libc_rand": True}, {"OMP_NUM_THREADS": "1"}, "rt_torus_m9_ao16"),
}


def channels(img):
    return np.stack([(img >> 16) & 255, (img >> 8) & 255, img & 255], -1).astype(np.int32)


def ao_statistics(only, index):
    for name, (model, mode, frames, kw, env, seeded_case) in STAT_CASES.items():
        if only and name not in only:
            continue
        if not pyport.have_ref(W, H, **kw):
            pyport.build_ref(W, H, **kw)
        imgs, _ = pyport.run_ref(pyport.model_path(model), mode, W, H, frames, env=env, **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{f"frame_{k}": v for k, v in imgs.items()})
        seeded = np.load(os.path.join(HERE, seeded_case + ".npz"))
        stats = {}
        for k, b in imgs.items():
            a = seeded[f"frame_{k}"]
            lit = (a != 0) | (b != 0)
            ca, cb = channels(a)[lit], channels(b)[lit]
            stats[str(k)] = {"lit_pixels": int(lit.sum()), "mean_abs_diff": round(float(np.abs(ca - cb).mean()), 3),
                             "max_abs_diff": int(np.abs(ca - cb).max()), "mean_level_seeded_stream": round(float(ca.mean()), 2),
                             "mean_level_libc_rand": round(float(cb.mean()), 2)}
        index[name] = {"model": model, "mode": mode, "frames": frames, "two_lights": False, "variant": kw, "width": W, "height": H,
                       "statistical": True, "seeded_case": seeded_case, "vs_seeded_stream": stats}
        print(name, stats)


def main():
    only = set(sys.argv[1:])
    index_path = os.path.join(HERE, "index.json")
    index = json.load(open(index_path)) if os.path.exists(index_path) else {}
    for name, (model, mode, frames, two, kw, env) in CASES.items():
        if only and name not in only:
            continue
        if not pyport.have_ref(W, H, **kw):
            pyport.build_ref(W, H, **kw)
        imgs, _ = pyport.run_ref(pyport.model_path(model), mode, W, H, frames, two_lights=two, env=env, **kw)
        arrs = {f"frame_{k}": v for k, v in imgs.items()}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
        index[name] = {"model": model, "mode": mode, "frames": frames, "two_lights": two, "variant": kw,
                       "width": W, "height": H,
                       "sha256": {str(k): hashlib.sha256(v.tobytes()).hexdigest() for k, v in imgs.items()}}
        print(name, {k: int((v != 0).sum()) for k, v in imgs.items()})
    full_size(only, index)
    ao_statistics(only, index)
    # the reference's own .bvh caches (byte-level pin of loader + BVH builder)
    bvh = {}
    for m in sorted(os.listdir(pyport.MODELS)):
        p = os.path.join(pyport.MODELS, m + "") if m.endswith(".bvh") else None
        if p:
            bvh[m[:-4]] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    if bvh:                      # merge: only the caches the reference wrote during THIS run are in the directory
        index.setdefault("_bvh_sha256", {}).update(bvh)
    json.dump(index, open(index_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
