"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the real reference.

Run in the authoring container (where /root/reference is mounted):

    python -m oracle.make_golden

It imports the unmodified reference under the import-time shims of
``oracle/ref_shim.py`` and records, for seeded inputs, what the reference's own
``Kiez(SklearnNN brute, hubness=...).fit().kneighbors()`` and
``kiez.analysis.hubness_score`` return.  It also converts the reference's own
known-answer fixtures for ``hubness_score`` (tests/nn_ind.npy,
tests/expected_k{2,5,10,50}_hub_scores.pkl -- data, not source) into one npz.
The committed npz files travel to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import os
import pickle
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

# (label, hubness name for the reference, kwargs)
HUBNESS_CASES = [
    ("no", None, {}),
    ("csls", "CSLS", {}),
    ("ls", "LocalScaling", {"method": "standard"}),
    ("nicdm", "LocalScaling", {"method": "nicdm"}),
    ("mp_gaussian", "MutualProximity", {"method": "normal"}),
    ("mp_empiric", "MutualProximity", {"method": "empiric"}),
    ("dsl", "DisSimLocal", {"squared": False}),
]


def datasets():
    """name -> (source, target_or_None, n_candidates, k, metrics)"""
    out = {}
    # the reference's session fixture, tests/conftest.py:5-11
    rng = np.random.RandomState(42)
    s, t = rng.rand(20, 5), rng.rand(50, 5)
    out["conftest"] = (s, t, 5, 3, ("euclidean", "cosine", "sqeuclidean"))
    out["conftest_single"] = (s, None, 5, 3, ("euclidean", "cosine"))
    # README example, README.md:61-71 (config C1)
    rng = np.random.RandomState(0)
    s, t = rng.rand(100, 50), rng.rand(100, 50)
    out["readme"] = (s, t, 10, 5, ("euclidean", "cosine"))
    out["readme_single"] = (s, None, 10, 5, ("euclidean",))
    # fp32-valued gaussian, ragged sizes, d not a multiple of 32
    rng = np.random.default_rng(7)
    s = rng.standard_normal((300, 40)).astype(np.float32).astype(np.float64)
    t = rng.standard_normal((421, 40)).astype(np.float32).astype(np.float64)
    out["gauss"] = (s, t, 20, 10, ("euclidean", "cosine"))
    # clustered ("hubby") unit vectors: near-ties and strong hubness
    rng = np.random.default_rng(11)
    cent = rng.standard_normal((12, 64))
    s = cent[rng.integers(0, 12, 257)] + 0.05 * rng.standard_normal((257, 64))
    t = cent[rng.integers(0, 12, 333)] + 0.05 * rng.standard_normal((333, 64))
    s = (s / np.linalg.norm(s, axis=1, keepdims=True)).astype(np.float32).astype(np.float64)
    t = (t / np.linalg.norm(t, axis=1, keepdims=True)).astype(np.float32).astype(np.float64)
    out["hubby"] = (s, t, 16, 10, ("euclidean",))
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    kiez = ref_shim.load_reference()
    from kiez.analysis import hubness_score

    for name, (s, t, c, k, metrics) in datasets().items():
        rec = {"source": s, "n_candidates": np.int64(c), "k": np.int64(k)}
        if t is not None:
            rec["target"] = t
        for metric in metrics:
            for label, hub, kw in HUBNESS_CASES:
                if label == "dsl" and metric == "cosine":
                    continue
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    d, i = ref_shim.reference_kneighbors(
                        s, t, metric=metric, hubness=hub, hubness_kwargs=kw,
                        n_candidates=c, k=k)
                rec[f"{metric}__{label}__dist"] = d
                rec[f"{metric}__{label}__ind"] = i
        np.savez_compressed(os.path.join(GOLDEN, f"kiez_{name}.npz"), **rec)
        print("wrote", name, len(rec), "arrays")

    # the reference's known-answer fixtures for hubness_score
    tdir = os.path.join(ref_shim.REFERENCE_ROOT, "tests")
    nn_ind = np.load(os.path.join(tdir, "nn_ind.npy"))
    rec = {"nn_ind": nn_ind.astype(np.int16), "target_samples": np.int64(1000)}
    for kk in (2, 5, 10, 50):
        with open(os.path.join(tdir, f"expected_k{kk}_hub_scores.pkl"), "rb") as fh:
            exp = pickle.load(fh)
        for key, val in exp.items():
            rec[f"k{kk}__{key}"] = np.asarray(val)
    # plus live reference outputs for k in {1,5,10} (test_estimation.py:77-98)
    for kk in (1, 3, 10):
        res = hubness_score(nn_ind, 1000, k=kk, return_value="all", store_k_occurrence=True)
        for key, val in res.items():
            rec[f"live_k{kk}__{key}"] = np.asarray(val)
    # the 5x2 toy (test_estimation.py:38-43)
    rec["toy_nn"] = np.array([[0, 2], [1, 0], [2, 0], [3, 1], [4, 0]])
    rec["toy_k_skewness"] = np.float64(0.9128709291752769)
    np.savez_compressed(os.path.join(GOLDEN, "hubness_score.npz"), **rec)
    print("wrote hubness_score", len(rec), "arrays")
    openea_golden()


def write_openea_dir(root, seed=3, n_rows=60, d=7):
    """A small OpenEA-layout directory (kiez/io/data_loading.py:35-40) with the awkward cases:
    interleaved rows of the two graphs, rows of no graph, an id past the end of the matrix, one
    entity listed for two rows, links for a subset of the entities."""
    rng = np.random.default_rng(seed)
    emb_dir, kg_dir = os.path.join(root, "emb"), os.path.join(root, "kg")
    os.makedirs(emb_dir, exist_ok=True)
    os.makedirs(kg_dir, exist_ok=True)
    emb = rng.standard_normal((n_rows, d)).astype(np.float32)
    np.save(os.path.join(emb_dir, "ent_embeds.npy"), emb)
    perm = rng.permutation(n_rows)
    rows1, rows2 = perm[:25], perm[25:52]            # 8 rows belong to neither graph
    with open(os.path.join(emb_dir, "kg1_ent_ids"), "w") as fh:
        for r in rows1:
            fh.write(f"http://kg1/e{r}\t{r}\n")
        fh.write(f"http://kg1/beyond\t{n_rows + 5}\n")           # no such row
        fh.write(f"http://kg1/e{rows1[0]}\t{perm[59]}\n")         # same entity, second row
    with open(os.path.join(emb_dir, "kg2_ent_ids"), "w") as fh:
        for r in rows2:
            fh.write(f"http://kg2/e{r}\t{r}\n")
    with open(os.path.join(kg_dir, "ent_links"), "w") as fh:
        for a, b in zip(rows1[1:21], rows2[:20]):
            fh.write(f"http://kg1/e{a}\thttp://kg2/e{b}\n")
    return emb_dir, kg_dir


def openea_golden():
    """tests/golden/openea_small/: the directory above + what the reference's own
    kiez.io.data_loading.from_openea returns for it."""
    import json

    from kiez.io.data_loading import from_openea

    root = os.path.join(GOLDEN, "openea_small")
    emb_dir, kg_dir = write_openea_dir(root)
    emb1, emb2, ids1, ids2, links = from_openea(emb_dir, kg_dir)
    np.savez_compressed(os.path.join(root, "expected.npz"), emb1=emb1, emb2=emb2)
    with open(os.path.join(root, "expected.json"), "w") as fh:
        json.dump({"kg1_ids": ids1, "kg2_ids": ids2,
                   "ent_links": {str(a): int(b) for a, b in links.items()}}, fh, indent=0,
                  sort_keys=True)
    print("wrote openea_small", emb1.shape, emb2.shape, len(links), "links")


if __name__ == "__main__":
    main()
