"""Harrell C-index -- restatement of lifelines.utils.concordance_index (third-party; NOT vendored in the reference).

TEST INFRASTRUCTURE.  Call sites: /root/reference/main.py:33,106-123 (`getCIndices`).  lifelines is unpinned in
/root/reference/requirements.txt:17.  Restated from the published lifelines algorithm:
pairs are (a, b) with b a death strictly earlier than a, or a death at the same time as a censored a; deaths at
equal times are not compared.  correct = #{P_b < P_a}, tied = #{P_b == P_a};  C = (correct + tied/2) / pairs in
float64;  pairs == 0 -> ZeroDivisionError("No admissable pairs in the dataset.");  NaN -> ValueError.
"""
import numpy as np


def _prep(event_times, predicted_scores, event_observed):
    t = np.asarray(event_times, dtype=np.float64).ravel()
    p = np.asarray(predicted_scores, dtype=np.float64).ravel()
    if event_observed is None:
        e = np.ones(t.shape[0], dtype=bool)
    else:
        e = np.asarray(event_observed, dtype=np.float64).ravel() != 0
    if t.shape != p.shape or t.shape != e.shape:
        raise ValueError("Observed events must be 1-dimensional of same length as event times")
    if np.isnan(t).any() or np.isnan(p).any():
        raise ValueError("NaNs detected in inputs, please correct or drop.")
    return t, p, e


def concordance_counts_bruteforce(event_times, predicted_scores, event_observed=None, weights=None):
    """O(N^2) definition.  With integer `weights` (multiplicities) gives the counts of the expanded multiset in
    which copies of one patient never pair with each other (SURVEY.md section 8a row 16 identity)."""
    t, p, e = _prep(event_times, predicted_scores, event_observed)
    n = t.shape[0]
    w = np.ones(n, dtype=np.int64) if weights is None else np.asarray(weights, dtype=np.int64)
    correct = tied = pairs = 0
    for a in range(n):
        if w[a] == 0:
            continue
        adm = e & ((t < t[a]) | ((t == t[a]) & (not e[a])))
        adm[a] = False
        wb = w[adm] * w[a]
        pairs += int(wb.sum())
        correct += int(wb[p[adm] < p[a]].sum())
        tied += int(wb[p[adm] == p[a]].sum())
    return correct, tied, pairs


def concordance_counts(event_times, predicted_scores, event_observed=None):
    """O(N log N) sweep: ascending time; at each distinct time deaths are counted against the pool then inserted,
    censored are counted against the pool (including the deaths at that time) and never inserted."""
    t, p, e = _prep(event_times, predicted_scores, event_observed)
    # rank-compress the scores so the pool is a Fenwick tree over score ranks
    uniq, rk = np.unique(p, return_inverse=True)
    m = uniq.shape[0]
    tree = np.zeros(m + 1, dtype=np.int64)

    def add(i):
        i += 1
        while i <= m:
            tree[i] += 1
            i += i & (-i)

    def prefix(i):  # number of pool entries with rank < i
        s = 0
        while i > 0:
            s += tree[i]
            i -= i & (-i)
        return int(s)

    order = np.lexsort((~e, t))  # by time, deaths before censored at equal time
    correct = tied = pairs = 0
    pool = 0
    i = 0
    n = t.shape[0]
    while i < n:
        j = i
        while j < n and t[order[j]] == t[order[i]]:
            j += 1
        grp = order[i:j]
        deaths = [g for g in grp if e[g]]
        cens = [g for g in grp if not e[g]]
        for g in deaths:
            lt = prefix(rk[g]); le = prefix(rk[g] + 1)
            pairs += pool; correct += lt; tied += le - lt
        for g in deaths:
            add(rk[g])
        pool += len(deaths)
        for g in cens:
            lt = prefix(rk[g]); le = prefix(rk[g] + 1)
            pairs += pool; correct += lt; tied += le - lt
        i = j
    return correct, tied, pairs


def concordance_index(event_times, predicted_scores, event_observed=None):
    correct, tied, pairs = concordance_counts(event_times, predicted_scores, event_observed)
    if pairs == 0:
        raise ZeroDivisionError("No admissable pairs in the dataset.")
    return (correct + tied / 2) / pairs


def getCIndices(preds, events, durations, num_classes=2):
    """/root/reference/main.py:106-123."""
    return [concordance_index(np.asarray(durations)[:, i], np.asarray(preds)[:, i], np.asarray(events)[:, i])
            for i in range(num_classes)]


def bootstrap_cindex(preds, events, durations, resample_indices, num_classes=2):
    """Idiomatic restatement of the bootstrap loop at /root/reference/main.py:768-887: patients are forwarded
    once, each resample is an index multiset; a resample with no admissible pair is skipped
    (ZeroDivisionError at main.py:856-858).  Returns (per-resample [R, C] float64 with NaN for skipped,
    mean [C], std ddof=0 [C])."""
    preds = np.asarray(preds); events = np.asarray(events); durations = np.asarray(durations)
    out = np.full((len(resample_indices), num_classes), np.nan)
    for r, idx in enumerate(resample_indices):
        try:
            out[r] = getCIndices(preds[idx], events[idx], durations[idx], num_classes)
        except ZeroDivisionError:
            pass
    ok = ~np.isnan(out).any(axis=1)
    return out, out[ok].mean(axis=0), out[ok].std(axis=0)
