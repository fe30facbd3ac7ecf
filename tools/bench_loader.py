"""CPU timing of the OpenEA row split: the reference's per-row loop (restated here exactly as
kiez/io/data_loading.py:23-32 walks the matrix; the reference itself is used instead when it is
mounted) against kiez_b200.io._split_emb (one vectorised gather).

    python tools/bench_loader.py [--rows 2000000 --d 256]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kiez_b200.io import _split_emb  # noqa: E402


def reference_split(emb, kg_ids):
    try:
        from oracle import ref_shim

        if ref_shim.reference_available():
            ref_shim.load_reference()
            from kiez.io.data_loading import _split_emb as ref

            return ref(emb, kg_ids), "reference (kiez.io.data_loading._split_emb)"
    except Exception:
        pass
    new_emb, new_ids, i = [], {}, 0          # restatement of the reference's loop
    for idx, e in enumerate(emb):
        if idx in kg_ids:
            new_ids[kg_ids[idx]] = i
            new_emb.append(e)
            i += 1
    return (np.array(new_emb), new_ids), "port of the reference's loop"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--d", type=int, default=256)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    emb = rng.standard_normal((args.rows, args.d), dtype=np.float32)
    rows = rng.permutation(args.rows)[: args.rows // 2]
    kg_ids = {int(r): f"e{r}" for r in rows}
    t0 = time.perf_counter()
    (want_emb, want_ids), what = reference_split(emb, kg_ids)
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    got_emb, got_ids = _split_emb(emb, kg_ids)
    t_new = time.perf_counter() - t0
    assert np.array_equal(got_emb, want_emb) and got_ids == want_ids
    print(f"{args.rows} x {args.d} fp32, {len(kg_ids)} rows selected: {what} {t_ref:.2f} s, "
          f"kiez_b200.io._split_emb {t_new:.2f} s ({t_ref / t_new:.1f}x), identical outputs")


if __name__ == "__main__":
    main()
