"""oracle/stats_oracle.py -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the per-contact statistics of the reference's AccumulatedContact, as plain Python
loops in the reference's own order of operations (repeated ``+= ns_per_frame`` and all):

  mean_score         PyContact/core/Biochemistry.py:179-186
  median_score       :188-192          (numpy.median)
  hbond_percentage   :150-158  over hbondFramesScan :215-224
  total_time         :160-167
  life_time          :194-213  (mean_life_time :169-172, median_life_time :174-177)

Pinned against the UNMODIFIED reference classes run on the reference's golden session
(tests/golden/make_stats_golden.py -> tests/golden/rpn11_ubq_stats_golden.npz, tests/test_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np


def life_time(scores, ns_per_frame, threshold):
    out, active, t = [], False, 0
    n = len(scores)
    for i, s in enumerate(scores):
        if not active and s > threshold:
            active = True
            t += ns_per_frame
        elif active and s > threshold:
            t += ns_per_frame
        elif active and s <= threshold:
            active = False
            out.append(t)
            t = 0
        if i == n - 1:
            out.append(t)
    return out


def key_statistics(score_matrix, hbond_matrix, ns_per_frame, threshold):
    """dict of float64 arrays, one value per key (row of the matrices)."""
    K = len(score_matrix)
    res = {k: np.zeros(K) for k in ("mean", "median", "hbond_percentage", "total_time", "mean_life_time",
                                    "median_life_time", "life_entries", "frames_above")}
    for k in range(K):
        row = [float(v) for v in score_matrix[k]]
        mean = 0
        for s in row:
            mean += s
        res["mean"][k] = mean / len(row)
        res["median"][k] = np.median(row)
        cnt = 0
        for h in hbond_matrix[k]:
            if h > 0:
                cnt += 1
        res["hbond_percentage"][k] = float(cnt) / float(len(row)) * 100
        t = 0
        above = 0
        for s in row:
            if s > threshold:
                t += ns_per_frame
                above += 1
        res["total_time"][k] = t
        lt = life_time(row, ns_per_frame, threshold)
        res["mean_life_time"][k] = np.mean(lt)
        res["median_life_time"][k] = np.median(lt)
        res["life_entries"][k] = len(lt)
        res["frames_above"][k] = above
    return res
