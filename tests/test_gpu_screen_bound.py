"""-m gpu: the error bound E of the screen's completeness proof (csrc/refine.cu) against the
keys the tcgen05 kernel ACTUALLY produces (accumulation order and truncation of the tensor core
included; tests/test_cpu_screen_bound.py can only emulate those).  For every (row, id) pair a
screen launch returns, |screen key - float64 key| must stay below the E the proof uses for that
row; the observed maximum of the ratio is printed.  One-direction form: the row lists
(kb2_knn_screen cand_idx / cand_key).  Dual-direction form: additionally the per-column emit
buffers, whose packed keys are what the column side of the proof (kb2_col_select's bound) sees.
>= 1e7 pairs per distribution and feature count."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _synth(kind, n, d, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    x = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)
    if kind == "gauss":
        return x
    if kind == "shifted":
        return x + 50.0
    if kind == "scaled":
        return x * 1e3 * (0.01 + 0.99 * torch.rand((n, 1), generator=g, device="cuda"))
    gc = torch.Generator(device="cuda")          # "hubby": bench.py's clustered unit vectors
    gc.manual_seed(12345)
    centres = torch.randn((40, d), generator=gc, device="cuda", dtype=torch.float32)
    assign = torch.randint(0, 40, (n,), generator=g, device="cuda")
    return torch.nn.functional.normalize(0.25 * x + centres[assign], dim=1)


def _proof_bound(algo, q, y):
    """E per query row, float64, exactly as refine_topk_kernel<CHECK> computes it (euclidean)."""
    up = 1.000001
    qn2 = q.key.double()
    ym2 = y.keymax.double()
    qn, ym = qn2.sqrt() * up, ym2.sqrt() * up
    dq, dym = q.err.double().sqrt() * up, y.errmax.double().sqrt() * up
    eps = algo._eps_acc(q.dpad)
    return 2.0 * (dq * ym + qn * (1.0 + 2.0 ** -11) * dym + eps * qn * ym) \
        + 4.76837158203125e-07 * (ym2 + 2.0 * qn * ym)


def _exact_keys(q_raw, y_raw, center, rows, ids):
    """||y-c||^2 - 2 <q-c, y-c> in float64 for the pairs (rows[i], ids[i, j])."""
    c = center.double()
    out = torch.empty(ids.shape, dtype=torch.float64, device=ids.device)
    step = max(1, (1 << 26) // (ids.shape[1] * q_raw.shape[1]))      # ~0.5 GB of gathered rows
    for lo in range(0, ids.shape[0], step):
        hi = min(ids.shape[0], lo + step)
        qc = q_raw[rows[lo:hi]].double() - c
        yc = y_raw[ids[lo:hi].clamp(min=0).reshape(-1)].double().view(hi - lo, ids.shape[1], -1) - c
        out[lo:hi] = (yc * yc).sum(dim=2) - 2.0 * torch.einsum("rd,rjd->rj", qc, yc)
    return out


def _unpack(ent):
    """packed (order-preserving key bits << 32 | id) -> (fp32 key, id): select.cuh entry_key/_col."""
    u = (ent >> 32) & 0xFFFFFFFF
    bits = torch.where((u >> 31) == 1, u ^ 0x80000000, u ^ 0xFFFFFFFF)
    bits = torch.where(bits >= (1 << 31), bits - (1 << 32), bits).to(torch.int32)
    return bits.view(torch.float32), (ent & 0xFFFFFFFF)


@pytest.mark.parametrize("kind", ["gauss", "shifted", "scaled", "hubby"])
@pytest.mark.parametrize("d", [128, 256])
def test_device_screen_keys_stay_within_the_proof_bound(kind, d):
    from kiez_b200 import B200

    algo = B200(n_candidates=10, precision="screen")
    cap = 32 if d == 256 else 64
    assert algo._lib.lib.kb2_screen_stages(d, cap, 0) > 0 and algo._lib.lib.kb2_screen_stages(d, cap, 1) > 0
    nq, ny = (10_500_000 // cap + 255) // 256 * 256, 40_000
    q_raw, y_raw = _synth(kind, nq, d, 1), _synth(kind, ny, d, 2)
    algo._center_vec = None
    y = algo._prepare(y_raw, cache=False)       # the first prepared matrix defines the centre
    q = algo._prepare(q_raw, cache=False)
    center = algo._center_vec
    E = _proof_bound(algo, q, y)

    # one-direction form: every row list
    cand, ckey, lists = algo._screen_search(q, y, cap)
    torch.cuda.synchronize()
    assert cand.shape == (nq, lists * cap) and int(cand.min()) >= 0
    rows = torch.arange(nq, device="cuda")
    exact = _exact_keys(q_raw, y_raw, center, rows, cand.long())
    ratio = ((ckey.double() - exact).abs() / E[:, None]).max().item()
    pairs = cand.numel()
    print(f"{kind} d={d} one-direction: {pairs} pairs, max |screen - exact| / E = {ratio:.3f}")
    assert pairs >= 10_000_000 and ratio < 1.0

    # dual-direction form: the row lists again + the column emit buffers
    col_cap = 1024
    n_s = 2048
    sample = q.take(torch.arange(n_s, device="cuda") * (nq // n_s))
    _i, s_key, s_lists = algo._screen_search(y, sample, cap)
    tau = s_key.view(ny, s_lists, cap)[:, :, cap - 1].amin(dim=1).contiguous()
    seg = q.rows(0, 64 * 256)
    col_cnt = torch.zeros(ny, dtype=torch.int32, device="cuda")
    col_buf = torch.zeros((ny, col_cap), dtype=torch.int64, device="cuda")
    cand2, ckey2, lists2 = algo._screen_search(seg, y, cap, dual=(tau, col_cnt, col_buf, col_cap, 0))
    torch.cuda.synchronize()
    exact2 = _exact_keys(q_raw, y_raw, center, rows[: seg.n], cand2.long())
    ratio2 = ((ckey2.double() - exact2).abs() / E[: seg.n, None]).max().item()
    # column side: exact column key = ||x-c||^2 - 2 <x-c, y-c>; its proof runs with the roles
    # swapped (query = the column, index = the rows)
    E_col = _proof_bound(algo, y, q)
    cnt = col_cnt.clamp(max=col_cap).long()
    valid = torch.arange(col_cap, device="cuda")[None, :] < cnt[:, None]
    keys, ids = _unpack(col_buf)
    ids = torch.where(valid, ids, torch.zeros_like(ids))
    cols = torch.arange(ny, device="cuda")
    exact_c = _exact_keys(y_raw, q_raw, center, cols, ids)
    err_c = torch.where(valid, (keys.double() - exact_c).abs(), torch.zeros_like(exact_c))
    ratio3 = (err_c / E_col[:, None]).max().item()
    n_emits = int(valid.sum())
    print(f"{kind} d={d} dual-direction: rows {ratio2:.3f}, columns {ratio3:.3f} "
          f"({cand2.numel()} + {n_emits} pairs)")
    assert n_emits > ny and ratio2 < 1.0 and ratio3 < 1.0
