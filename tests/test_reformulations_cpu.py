"""Reformulations the CUDA path relies on, restated in numpy and checked against the oracle in float64:
they must be exact in real arithmetic, whatever the kernels then do with them."""
import numpy as np
import pytest

from oracle import message_passing_oracle as mo


@pytest.mark.parametrize("H", [32, 96, 128])
def test_gru_update_as_one_contraction_with_a_gate_epilogue(H):
    """gemm_tc.cu: pack_gru_weights_kernel + the GRU epilogue.  [agg | h] x W_gru [2H, 4H], whose 128-column tile t holds for
    the hidden units 32t..32t+31 the pre-activations z | r | x_h | h_h (x_h has no recurrent rows, h_h no input rows), then
    z = sigmoid, r = sigmoid, hh = tanh(x_h + r * h_h), h' = z h + (1 - z) hh  ==  Keras GRUCell(reset_after=True)."""
    rng = np.random.default_rng(H)
    V = 50
    agg, h = rng.standard_normal((V, H)), rng.standard_normal((V, H))
    K, U = rng.standard_normal((H, 3 * H)) * 0.3, rng.standard_normal((H, 3 * H)) * 0.3
    bias = rng.standard_normal((2, 3 * H)) * 0.1
    ref = mo.gru_cell_forward(agg, h, K, U, bias)

    N = 4 * H
    W = np.zeros((2 * H, N))
    b = np.zeros(N)
    for n in range(N):
        t, gate, u = n // 128, (n % 128) // 32, 32 * (n // 128) + n % 32
        if gate < 3:
            W[:H, n] = K[:, gate * H + u]
        if gate != 2:
            W[H:, n] = U[:, (2 if gate == 3 else gate) * H + u]
        b[n] = (bias[0, u] + bias[1, u] if gate == 0 else bias[0, H + u] + bias[1, H + u] if gate == 1
                else bias[0, 2 * H + u] if gate == 2 else bias[1, 2 * H + u])
    pre = np.concatenate([agg, h], axis=1) @ W + b
    out = np.empty((V, H))
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    for t in range(H // 32):
        z, r, xh, hh_ = (pre[:, 128 * t + 32 * g: 128 * t + 32 * g + 32] for g in range(4))
        z, r = sig(z), sig(r)
        cand = np.tanh(xh + r * hh_)
        out[:, 32 * t: 32 * t + 32] = z * h[:, 32 * t: 32 * t + 32] + (1.0 - z) * cand
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12)


def test_rgat_score_halves_are_row_dot_products_of_the_projection():
    """gemm_tc.cu epilogue (epi.score_src): with P_l = h W_l, the per-edge logit a_l[k] . [P_l[u,k,:] || P_l[v,k,:]] of
    rgat.py:111-121 is s_src[u,l,k] + s_tgt[v,l,k], each a dot product over ONE row of P - so both come out of the
    projection's epilogue."""
    rng = np.random.default_rng(3)
    V, D, H, Kh, L = 40, 16, 64, 4, 3
    d = H // Kh
    h = rng.standard_normal((V, D))
    W = [rng.standard_normal((D, H)) for _ in range(L)]
    att = [rng.standard_normal((Kh, 2 * d)) for _ in range(L)]
    for l in range(L):
        P = (h @ W[l]).reshape(V, Kh, d)
        s_src = np.einsum("vkd,kd->vk", P, att[l][:, :d])
        s_tgt = np.einsum("vkd,kd->vk", P, att[l][:, d:])
        u, v = rng.integers(0, V, 200), rng.integers(0, V, 200)
        literal = np.einsum("ekd,kd->ek", np.concatenate([P[u], P[v]], axis=2), att[l])
        np.testing.assert_allclose(s_src[u] + s_tgt[v], literal, rtol=0, atol=1e-12)


def test_online_softmax_with_one_exponential_per_edge():
    """rgat.cu PASS 2: either the running maximum grows (this edge weighs exp(0) = 1, the sums so far are rescaled) or it
    stays (the edge weighs exp(score - m)): equal to the two-pass softmax-weighted sum."""
    rng = np.random.default_rng(5)
    for n in (1, 2, 17, 300):
        score, x = rng.standard_normal(n) * 4, rng.standard_normal((n, 8))
        m, den, acc = -np.finfo(np.float64).max, 0.0, np.zeros(8)
        for s, row in zip(score, x):
            if s > m:
                rescale = np.exp(m - s)
                m, den, acc = s, den * rescale + 1.0, acc * rescale + row
            else:
                w = np.exp(s - m)
                den, acc = den + w, acc + w * row
        w2 = np.exp(score - score.max())
        np.testing.assert_allclose(acc / den, (w2[:, None] * x).sum(0) / w2.sum(), rtol=1e-12, atol=1e-12)


def test_quad_transpose_of_the_epilogue_stores():
    """sm100_ptx.cuh quad_transpose_f4: two xor-shuffle rounds leave lane t of a quad with float4 number t of the quad's
    four rows (simulated lane by lane)."""
    A = [[(row, piece) for piece in range(4)] for row in range(4)]
    shfl = lambda vals, m: [vals[t ^ m] for t in range(4)]
    b0, b1 = [t & 1 for t in range(4)], [(t >> 1) & 1 for t in range(4)]
    r0 = shfl([A[t][0] if b0[t] else A[t][1] for t in range(4)], 1)
    r1 = shfl([A[t][2] if b0[t] else A[t][3] for t in range(4)], 1)
    B = [[r0[t] if b0[t] else A[t][0], A[t][1] if b0[t] else r0[t], r1[t] if b0[t] else A[t][2],
          A[t][3] if b0[t] else r1[t]] for t in range(4)]
    s0 = shfl([B[t][0] if b1[t] else B[t][2] for t in range(4)], 2)
    s1 = shfl([B[t][1] if b1[t] else B[t][3] for t in range(4)], 2)
    C = [[s0[t] if b1[t] else B[t][0], s1[t] if b1[t] else B[t][1], B[t][2] if b1[t] else s0[t],
          B[t][3] if b1[t] else s1[t]] for t in range(4)]
    assert all(C[t][k] == (k, t) for t in range(4) for k in range(4))
