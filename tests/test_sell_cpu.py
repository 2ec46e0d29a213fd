"""Host half of the SELL-P build (csrc/spmv_b200.cu: sell_plan_host, exposed as
spmv_b200_sell_plan) against the oracle's independent numpy restatement of the layout rule
(oracle.sellp_layout): row order inside every window and panel, slice widths and offsets.
The device half (row counts per panel, slice fill) is covered by the -m gpu tests."""
import numpy as np

from conftest import random_csr


def _counts(M, N, IRP, JA, K, max_row):
    pc = [N * p // K for p in range(K + 1)]
    lens = np.diff(IRP)
    counts = np.zeros((K, M), np.int32)
    for r in range(M):
        cols = JA[IRP[r]:IRP[r + 1]]
        for p in range(K):
            if lens[r] > max_row:
                counts[p, r] = -1
            elif K == 1:
                counts[p, r] = len(cols)
            else:
                counts[p, r] = int(((cols >= pc[p]) & (cols < pc[p + 1])).sum())
    return counts


def test_sell_plan_matches_oracle(sp, O):
    L = sp._lib
    rng = np.random.default_rng(1)
    for trial in range(40):
        M = int(rng.integers(1, 900))
        N = int(rng.integers(1, 400))
        K = int(rng.integers(1, 6))
        sigma = int(rng.choice([32, 64, 256, 1024]))
        max_row = int(rng.choice([8, 20, 4096]))
        IRP, JA, AS = random_csr(rng, M, N, 14, empty_frac=0.3)
        if rng.random() < 0.5 and M > 2:       # a hub row
            extra = rng.integers(0, N, 70).astype(np.int32)
            r = int(rng.integers(0, M))
            JA = np.concatenate([JA[:IRP[r]], extra, JA[IRP[r]:]])
            AS = np.concatenate([AS[:IRP[r]], rng.uniform(-1, 1, 70), AS[IRP[r]:]])
            IRP = IRP.copy()
            IRP[r + 1:] += 70
        w_soff, w_perm, _, _ = O.sellp_layout(M, N, IRP, JA, AS, K, sigma, max_row)
        counts = _counts(M, N, IRP, JA, K, max_row)
        S = (M + 31) // 32
        perm = np.zeros(K * S * 32, np.int32)
        soff = np.zeros(K * (S + 1), np.int64)
        rc = L.b200.spmv_b200_sell_plan(counts.ctypes.data_as(L.c_ip), M, K, sigma, perm.ctypes.data_as(L.c_ip),
                                        soff.ctypes.data_as(L.c_i64p))
        assert rc == 0
        assert np.array_equal(perm.reshape(K, -1), w_perm), (trial, M, K, sigma, max_row)
        assert np.array_equal(soff.reshape(K, -1), w_soff), (trial, M, K, sigma, max_row)
        # every row that belongs to a slice appears exactly once per panel
        for p in range(K):
            rows = perm.reshape(K, -1)[p]
            rows = rows[rows >= 0]
            assert len(rows) == len(set(rows.tolist())) == int((counts[p] >= 0).sum())


def test_sell_plan_rejects_bad_arguments(sp):
    L = sp._lib
    c = np.zeros(10, np.int32)
    perm = np.zeros(32, np.int32)
    soff = np.zeros(2, np.int64)
    for K, sigma in ((0, 32), (65, 32), (1, 31), (1, 48)):
        assert L.b200.spmv_b200_sell_plan(c.ctypes.data_as(L.c_ip), 10, K, sigma, perm.ctypes.data_as(L.c_ip),
                                          soff.ctypes.data_as(L.c_i64p)) != 0


def test_virtual_row_plan_matches_oracle(sp, O):
    """Ragged matrices: rows longer than `chunk` are cut into virtual rows; destinations, slice
    offsets and the split-row bookkeeping equal the oracle's restatement (oracle.sellv_layout)."""
    L = sp._lib
    rng = np.random.default_rng(4)
    for trial in range(30):
        M = int(rng.integers(1, 600))
        N = 200
        lens = rng.integers(0, 25, M)
        for _ in range(int(rng.integers(0, 4))):
            lens[rng.integers(0, M)] = int(rng.integers(60, 1500))       # hub rows
        IRP = np.zeros(M + 1, np.int64)
        IRP[1:] = np.cumsum(lens)
        JA = rng.integers(0, N, int(IRP[-1])).astype(np.int32)
        AS = rng.uniform(-1, 1, int(IRP[-1]))
        chunk = int(rng.choice([4, 8, 64, 200]))
        sigma = int(rng.choice([32, 256, 16384]))
        w_soff, w_dest, _, _, w_srow, w_sfirst = O.sellv_layout(M, IRP, JA, AS, sigma, chunk)
        sizes = np.zeros(4, np.int64)
        assert L.b200.spmv_b200_sell_plan_vrows(IRP.ctypes.data_as(L.c_i64p), M, chunk, sigma,
                                                sizes.ctypes.data_as(L.c_i64p), None, None) == 0
        V, S, n_split, n_pieces = (int(v) for v in sizes)
        assert (S, n_split, n_pieces) == (w_soff.shape[1] - 1, len(w_srow), int(w_sfirst[-1]))
        assert V == int(np.where(lens > chunk, np.ceil(lens / chunk), 1).sum())
        dest = np.zeros(S * 32, np.int32)
        soff = np.zeros(S + 1, np.int64)
        assert L.b200.spmv_b200_sell_plan_vrows(IRP.ctypes.data_as(L.c_i64p), M, chunk, sigma,
                                                sizes.ctypes.data_as(L.c_i64p), dest.ctypes.data_as(L.c_ip),
                                                soff.ctypes.data_as(L.c_i64p)) == 0
        assert np.array_equal(dest, w_dest[0]), (trial, M, chunk, sigma)
        assert np.array_equal(soff, w_soff[0]), (trial, M, chunk, sigma)
        assert ((soff[1:] - soff[:-1]) // 32).max(initial=0) <= chunk     # no warp walks more than `chunk` steps
