"""Developer tool (run on a GPU box): compare the CUDA path with the oracle on a small
batch and, on a mismatch, bisect the op prefix that first diverges.

    python tests/gpu_debug.py [n_utts] [speed]
"""
from __future__ import annotations

import sys
import time

import numpy as np

import harness as H

gpu = H.importlib.import_module("2026-simple-c-tts_b200.gpu")


def prefix_plan(plan, u: int, k: int, speed: float):
    ops = plan.utt_ops(u)[:k].copy()
    return H.front.BatchPlan(np.array([0, len(ops)], np.uint32), np.array([speed], np.float32), ops)


def main() -> int:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    speed = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    raw = H.synthetic_db()
    cfg = H.shipped_config()
    fr = H.front.Front(raw, cfg, H.NORM_CSV)
    prm = fr.params()
    orc = H.Oracle(raw)
    texts = ["olá mundo", "a", "", "Olá, mundo! Como vai você? 123 casas."] + H.corpus.batch(max(n - 4, 0), seed=5)
    plan = fr.plan(texts, [speed] * len(texts))
    g = gpu.GpuSynth(raw, 0)
    t0 = time.time()
    outs = g.synth_list(plan, prm)
    print(f"gpu synth_batch {len(texts)} utts: {time.time() - t0:.3f}s")
    bad = 0
    for u in range(plan.n_utts):
        o, st = orc.synth(prm, plan.utt_ops(u), speed)
        r = outs[u]
        if len(o) == len(r) and np.array_equal(o, r):
            continue
        bad += 1
        if len(o) == len(r):
            d = np.nonzero(o != r)[0]
            print(f"utt {u}: {len(d)} diffs of {len(o)}, first {d[:6]}, maxabs {np.abs(o.astype(int) - r.astype(int)).max()}")
        else:
            print(f"utt {u}: LEN oracle {len(o)} gpu {len(r)}")
        if bad <= 2:
            ops = plan.utt_ops(u)
            lo, hi = 0, len(ops)
            # smallest prefix whose output differs (assembly only: speed 1.0)
            while lo < hi:
                mid = (lo + hi) // 2
                pp = prefix_plan(plan, u, mid, 1.0)
                go = g.synth_list(pp, prm)[0]
                oo, _ = orc.synth(prm, pp.utt_ops(0), 1.0)
                same = len(go) == len(oo) and np.array_equal(go, oo)
                if same:
                    lo = mid + 1
                else:
                    hi = mid
            if lo <= len(ops) and lo > 0:
                op = ops[lo - 1]
                pp = prefix_plan(plan, u, lo, 1.0)
                go = g.synth_list(pp, prm)[0]
                oo, _ = orc.synth(prm, pp.utt_ops(0), 1.0)
                msg = f"  first diverging prefix: {lo} ops; last op kind={op['kind']} flags={op['flags']} a={op['a']} b={op['b']}"
                if len(go) == len(oo):
                    d = np.nonzero(go != oo)[0]
                    msg += f"; {len(d)} diffs in [{d.min()},{d.max()}] of {len(oo)}; gpu {go[d[:4]]} oracle {oo[d[:4]]}"
                else:
                    msg += f"; len gpu {len(go)} oracle {len(oo)}"
                print(msg)
            else:
                print("  assembly prefix identical at every length: divergence is in the stretch stage")
    print(f"mismatching utterances: {bad} of {plan.n_utts}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
