"""Loading entity embeddings in the OpenEA layout -- the data format on the input side of the
hot path.

Mirrors ``kiez.io.data_loading`` (kiez/io/data_loading.py:8-99) name for name and result for
result: ``from_openea(emb_dir_path, kg_path)`` reads ``ent_embeds.npy``, ``kg1_ent_ids``,
``kg2_ent_ids`` (``entity<TAB>row`` lines) and ``ent_links`` (``entity1<TAB>entity2`` lines) and
returns ``(emb1, emb2, kg1_ids_new, kg2_ids_new, ent_links_new)``.

What differs is how the rows are split.  The reference walks the embedding matrix in a Python
loop and appends row by row (``_split_emb``, data_loading.py:23-32: ~1 us per row plus one
``np.array`` of a list of rows, seconds for the 10^6-row matrices whose kNN takes under a
second here); this module selects the rows with one vectorised gather -- on the host, or, with
``device=...``, on the GPU after a single upload of the shared matrix, so that the two matrices
``Kiez.fit`` needs are already resident.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np

try:  # torch is plumbing: only the optional device path needs it
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _read_kg_ids(path) -> Dict[int, str]:
    """``entity<TAB>row`` lines -> {row: entity} (data_loading.py:8-13; a repeated row keeps its
    last entity, like the reference's dict comprehension)."""
    out: Dict[int, str] = {}
    with open(path) as in_file:
        for line in in_file:
            fields = line.strip().split("\t")
            out[int(fields[1])] = fields[0]
    return out


def _read_ent_links(path) -> Dict[str, str]:
    """``entity1<TAB>entity2`` lines -> {entity1: entity2} (data_loading.py:16-20)."""
    out: Dict[str, str] = {}
    with open(path) as in_file:
        for line in in_file:
            fields = line.strip().split("\t")
            out[fields[0]] = fields[1]
    return out


def _select_rows(n_rows: int, kg_ids: Dict[int, str]):
    """Rows of the shared matrix that belong to one knowledge graph, in matrix order, and
    {entity: position among those rows} -- what ``_split_emb`` (data_loading.py:23-32) builds
    while it walks the matrix.  Rows listed in ``kg_ids`` but absent from the matrix are skipped,
    an entity listed for several rows ends at its last one (the reference overwrites
    ``new_ids[entity]`` as it goes)."""
    if kg_ids:
        rows = np.fromiter(kg_ids.keys(), dtype=np.int64, count=len(kg_ids))
        rows = np.sort(rows[(rows >= 0) & (rows < n_rows)])
    else:
        rows = np.empty((0,), dtype=np.int64)
    new_ids = {kg_ids[r]: i for i, r in enumerate(rows.tolist())}   # tolist(): plain ints
    return rows, new_ids


def _split_emb(emb, kg_ids: Dict[int, str]):
    """(rows of ``emb`` listed in ``kg_ids``, {entity: new row}); one gather instead of the
    reference's per-row loop.  Like the reference, no matching row gives ``np.array([])``."""
    rows, new_ids = _select_rows(len(emb), kg_ids)
    if rows.size == 0:
        return np.array([]), new_ids
    emb = np.asarray(emb)
    if torch is not None and emb.ndim == 2 and emb.dtype in (np.float16, np.float32, np.float64) \
            and emb.flags.c_contiguous and emb.flags.writeable:
        # multi-threaded row gather (numpy's fancy indexing copies 1 KB rows on one core)
        return torch.from_numpy(emb).index_select(0, torch.from_numpy(rows)).numpy(), new_ids
    return emb[rows], new_ids


def _read_openea_files(emb_dir_path, kg_path):
    """data_loading.py:35-40."""
    emb = np.load(os.path.join(emb_dir_path, "ent_embeds.npy"))
    kg1_ids = _read_kg_ids(os.path.join(emb_dir_path, "kg1_ent_ids"))
    kg2_ids = _read_kg_ids(os.path.join(emb_dir_path, "kg2_ent_ids"))
    ent_links = _read_ent_links(os.path.join(kg_path, "ent_links"))
    return emb, kg1_ids, kg2_ids, ent_links


def _seperate_common_embedding(
    emb: np.ndarray,
    kg1_ids: Dict[int, str],
    kg2_ids: Dict[int, str],
    ent_links: Dict[str, str],
    device: Optional[object] = None,
) -> Tuple[object, object, Dict[str, int], Dict[str, int], Dict[int, int]]:
    """Separate the shared embedding matrix into one matrix per knowledge graph
    (data_loading.py:43-72; the reference's spelling of the name is kept).

    Returns ``emb1, emb2, kg1_ids_new, kg2_ids_new, ent_links_new``: the rows of each graph in
    matrix order, {entity: row of its matrix} per graph, and the gold links as
    {row of emb1: row of emb2}.  A link that names an entity without an embedding row raises
    ``KeyError``, as in the reference.  With ``device`` the two matrices are torch tensors on
    that device (one upload of ``emb``, two on-device gathers)."""
    if device is None:
        emb1, kg1_ids_new = _split_emb(emb, kg1_ids)
        emb2, kg2_ids_new = _split_emb(emb, kg2_ids)
    else:
        if torch is None:
            raise ImportError("from_openea(device=...) needs PyTorch")
        rows1, kg1_ids_new = _select_rows(len(emb), kg1_ids)
        rows2, kg2_ids_new = _select_rows(len(emb), kg2_ids)
        dev = torch.device(device)
        emb_dev = torch.as_tensor(np.ascontiguousarray(emb)).to(dev, non_blocking=True)
        emb1 = emb_dev.index_select(0, torch.from_numpy(rows1).to(dev))
        emb2 = emb_dev.index_select(0, torch.from_numpy(rows2).to(dev))
    ent_links_new = {kg1_ids_new[e1]: kg2_ids_new[e2] for e1, e2 in ent_links.items()}
    return emb1, emb2, kg1_ids_new, kg2_ids_new, ent_links_new


def from_openea(emb_dir_path: str, kg_path: str, device: Optional[object] = None):
    """Load OpenEA-type data (data_loading.py:75-99; dataset layout:
    https://github.com/nju-websoft/OpenEA#dataset-description).

    Parameters
    ----------
    emb_dir_path: folder with ``ent_embeds.npy``, ``kg1_ent_ids``, ``kg2_ent_ids``
    kg_path: folder with ``ent_links``
    device: optional torch device -- return the two matrices as tensors on it

    Returns
    -------
    emb1, emb2, kg1_ids_new, kg2_ids_new, ent_links_new
    """
    return _seperate_common_embedding(*_read_openea_files(emb_dir_path, kg_path), device=device)
