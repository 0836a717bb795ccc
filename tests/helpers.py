"""Tie-aware comparisons shared by the parity tests."""
import numpy as np


def assert_topk_matches(D, I, D_ref, I_ref, score_tol, largest, what=""):
    """ids must be identical except where the reference's own scores are within score_tol of each
    other (documented near-tie exemption); scores must agree to score_tol."""
    assert D.shape == D_ref.shape and I.shape == I_ref.shape, f"{what}: shape"
    nq, k = I.shape
    bad = []
    for q in range(nq):
        ref_ids = I_ref[q]
        got_ids = I[q]
        assert (got_ids >= 0).sum() == (ref_ids >= 0).sum(), f"{what}: q{q} padding differs"
        valid = ref_ids >= 0
        np.testing.assert_allclose(D[q][valid], D_ref[q][valid], rtol=0, atol=score_tol,
                                   err_msg=f"{what}: q{q} scores")
        for j in range(k):
            if got_ids[j] == ref_ids[j]:
                continue
            # a differing id is acceptable only inside a run of reference scores within tol
            s = D_ref[q][j]
            close = np.abs(D_ref[q] - s) <= 2 * score_tol
            pos = np.nonzero(ref_ids == got_ids[j])[0]
            ok = (pos.size > 0 and close[pos[0]]) or (j == k - 1 or close[k - 1])
            if not ok:
                bad.append((q, j, int(got_ids[j]), int(ref_ids[j])))
    assert not bad, f"{what}: id mismatches outside near-ties: {bad[:10]}"


def recall_at_k(I, I_ref):
    hits = 0
    tot = 0
    for a, b in zip(I, I_ref):
        b = b[b >= 0]
        hits += len(set(a.tolist()) & set(b.tolist()))
        tot += len(b)
    return hits / max(tot, 1)
