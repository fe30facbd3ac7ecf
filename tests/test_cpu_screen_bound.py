"""not gpu: the error bound E behind the screen's completeness proof (csrc/refine.cu,
refine_topk_kernel<CHECK>; B200._eps_acc, kb2_split_error_terms) against a numpy emulation of the screen's arithmetic:
operands centred in fp32 and rounded to TF32 (cvt.rna: 10 mantissa bits, ties away, prep.cu),
products accumulated in fp32, key terms rounded to fp32.  If |screen key - exact key| <= E for
every (query, index) pair, "exact k-th key < tau - E" proves that no row outside the proposal
list can be among the k nearest -- the claim the GPU tests rely on.  The emulation cannot know
the tensor core's internal summation order, so the accumulation is checked in two orders."""
import numpy as np
import pytest


def to_tf32(x):
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) to 10 explicit mantissa bits."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def eps_acc(dpad):
    """B200._eps_acc (kiez_b200/neighbors.py)."""
    return dpad * 2.0 ** -22 + 2.0 ** -21


def split_error_terms(w):
    """kb2_split_error_terms (prep.cu): upper bound of ||w - hi||^2 per row from lo = tf32(w - hi)."""
    lo = to_tf32((w - to_tf32(w)).astype(np.float32))
    acc = (lo.astype(np.float64) ** 2).sum(axis=1) * (1.0 + 2.0 ** -9) * (1.0 + 2.0 ** -20)
    out = acc.astype(np.float32)
    return np.where(out.astype(np.float64) < acc, np.nextafter(out, np.float32(np.inf)), out)


def proof_bound(qn2, ym2, q_err, y_err_max, dpad):
    """E of refine.cu from the fp32 inputs the kernel reads."""
    up = 1.000001
    qn, ym = np.sqrt(qn2) * up, np.sqrt(ym2) * up
    dq, dym = np.sqrt(q_err.astype(np.float64)) * up, np.sqrt(float(y_err_max)) * up
    return (2.0 * (dq * ym + qn * (1.0 + 2.0 ** -11) * dym + eps_acc(dpad) * qn * ym)
            + 4.76837158203125e-07 * (ym2 + 2.0 * qn * ym))


def screen_keys(q, y, center, order):
    """key = ||y-c||^2 (fp32) - 2 <tf32(q-c), tf32(y-c)> accumulated in fp32."""
    qc = (q - center).astype(np.float32)                 # fp32 subtraction, as prep.cu
    yc = (y - center).astype(np.float32)
    y_key = (yc.astype(np.float64) ** 2).sum(axis=1).astype(np.float32)
    q_key = (qc.astype(np.float64) ** 2).sum(axis=1).astype(np.float32)
    qh, yh = to_tf32(qc), to_tf32(yc)
    if order == "blas":
        dot = qh @ yh.T                                   # fp32 accumulate, library order
    else:                                                 # K blocks of 8, sequential fp32 adds
        dot = np.zeros((q.shape[0], y.shape[0]), dtype=np.float32)
        for k0 in range(0, q.shape[1], 8):
            dot = (dot + (qh[:, k0:k0 + 8].astype(np.float64)
                          @ yh[:, k0:k0 + 8].T.astype(np.float64)).astype(np.float32)).astype(np.float32)
    key = (y_key[None, :] + (np.float32(-2.0) * dot)).astype(np.float32)
    return key, q_key, y_key, split_error_terms(qc), split_error_terms(yc)


def data(kind, n, d, rng):
    if kind == "gauss":
        return rng.standard_normal((n, d))
    if kind == "shifted":
        return 50.0 + rng.standard_normal((n, d))
    if kind == "scaled":
        return 1e3 * rng.standard_normal((n, d)) * rng.uniform(0.01, 1.0, (n, 1))
    cent = np.random.default_rng(7).standard_normal((6, d))     # tight clusters of unit vectors
    x = cent[rng.integers(0, 6, n)] + 0.02 * rng.standard_normal((n, d))
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("kind", ["gauss", "shifted", "scaled", "clusters"])
@pytest.mark.parametrize("d", [32, 256])
@pytest.mark.parametrize("order", ["blas", "k8"])
def test_screen_key_error_is_within_the_proof_bound(kind, d, order):
    rng = np.random.default_rng(d + len(kind))
    q = data(kind, 96, d, rng).astype(np.float32)
    y = data(kind, 700, d, rng).astype(np.float32)
    center = y.mean(axis=0, dtype=np.float64).astype(np.float32)       # B200._center_vec
    key, q_key, y_key, q_err, y_err = screen_keys(q, y, center, order)
    q64, y64 = q.astype(np.float64), y.astype(np.float64)
    d2 = ((q64[:, None, :] - y64[None, :, :]) ** 2).sum(axis=2)        # what the exact finish computes
    exact_key = d2 - q_key.astype(np.float64)[:, None]
    # E as in refine.cu (euclidean branch): norms inflated by 1e-6, key_max = max ||y-c||^2
    E = proof_bound(q_key.astype(np.float64), float(y_key.max()), q_err, y_err.max(), d)
    # the proof also gives away 2.4e-7 * ||q-c||^2 for the fp32 rounding of q_key
    slack = E + 2.4e-7 * q_key.astype(np.float64)
    err = np.abs(key.astype(np.float64) - exact_key)
    ratio = (err / slack[:, None]).max()
    assert ratio <= 1.0, f"{kind} d={d}: |screen - exact| exceeds E by {ratio:.3f}x"
    # and the bound is not vacuous: within ~2 orders of magnitude of the observed error
    assert ratio > 1e-3, f"{kind} d={d}: bound {1 / ratio:.0f}x looser than any observed error"
    # the measured-rounding-error bound is what lets clustered data pass the proof: at least 2x
    # tighter than the operand-independent 2^-10 ||q|| ||y|| bound it replaced
    qn = np.sqrt(q_key.astype(np.float64))
    old = 2.0 * (2.0 ** -10 + d * 2.0 ** -23) * qn * np.sqrt(float(y_key.max()))
    assert np.median(E / old) < 0.55 and (E / old).max() < 0.75
    print(f'{kind} d={d} {order}: max err/E {ratio:.3f}, E/old median {np.median(E / old):.3f}')


def test_tf32_rounding_emulation():
    x = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 + 2.0 ** -20, -3.0 - 2.0 ** -10,
                  1e-30, 0.0], dtype=np.float32)
    got = to_tf32(x)
    want = np.array([1.0, 1.0 + 2.0 ** -10, 1.0 + 2.0 ** -10, -3.0 - 2.0 ** -9, 1e-30, 0.0],
                    dtype=np.float32)
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2] and got[5] == 0.0
    assert got[3] == np.float32(-3.0 - 2.0 ** -9) or got[3] == np.float32(-3.0)   # tie: away
    assert abs(float(got[4]) - 1e-30) <= 1e-30 * 2.0 ** -11
    rel = np.abs(to_tf32(np.float32(np.random.default_rng(0).standard_normal(10000))) /
                 np.float32(np.random.default_rng(0).standard_normal(10000)) - 1.0)
    assert rel.max() <= 2.0 ** -11 * (1 + 1e-6)


@pytest.mark.parametrize("kind", ["gauss", "shifted", "clusters"])
@pytest.mark.parametrize("d", [32, 256])
def test_cosine_screen_key_error_is_within_the_proof_bound(kind, d):
    """Cosine branch of the proof (refine.cu): rows are L2-normalised in fp32 before the TF32
    rounding (prep.cu, normalize = 1), key = -2 <q^, y^>, exact key = 2 (cosine distance - 1),
    E as in the euclidean branch with unit norms, + 1e-6."""
    rng = np.random.default_rng(3 * d + len(kind))
    q = data(kind, 96, d, rng).astype(np.float32)
    y = data(kind, 700, d, rng).astype(np.float32)

    def prep(x):
        n2 = (x.astype(np.float64) ** 2).sum(axis=1)
        scale = (1.0 / np.sqrt(n2)).astype(np.float32)
        w = (x * scale[:, None]).astype(np.float32)
        return to_tf32(w), split_error_terms(w)

    dot = np.zeros((q.shape[0], y.shape[0]), dtype=np.float32)
    (qh, q_err), (yh, y_err) = prep(q), prep(y)
    for k0 in range(0, d, 8):
        dot = (dot + (qh[:, k0:k0 + 8].astype(np.float64)
                      @ yh[:, k0:k0 + 8].T.astype(np.float64)).astype(np.float32)).astype(np.float32)
    key = (np.float32(-2.0) * dot).astype(np.float64)
    q64, y64 = q.astype(np.float64), y.astype(np.float64)
    cos = (q64 @ y64.T) / np.outer(np.linalg.norm(q64, axis=1), np.linalg.norm(y64, axis=1))
    exact_key = 2.0 * ((1.0 - cos) - 1.0)
    E = proof_bound(np.ones(q.shape[0]), 1.0, q_err, y_err.max(), d) + 1e-6
    ratio = (np.abs(key - exact_key) / E[:, None]).max()
    assert ratio <= 1.0, f"{kind} d={d}: cosine screen error exceeds E by {ratio:.3f}x"
    assert ratio > 1e-3
