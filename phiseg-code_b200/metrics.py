"""Host side of the on-device validation metrics (csrc/metrics.cu): turns the label-intersection counts and the
cross-entropy-map moments the kernels produce into the scalar scores of the reference's validation loop
(phiseg_model.py:558-640): generalised energy distance (utils.py:270-320), variance NCC (utils.py:103-118,323-362) and
per-label Dice (medpy dc as used at phiseg_model.py:596-606).  Only a few hundred integers / doubles cross the bus."""
import numpy as np


def _iou(inter, ca, cb):
    """per-label IoU with the reference's conventions (utils.py:281-287): both masks empty -> 1, one empty -> 0"""
    inter = np.asarray(inter, np.float64)
    ca = np.asarray(ca, np.float64)
    cb = np.asarray(cb, np.float64)
    union = ca + cb - inter
    both_empty = (ca == 0) & (cb == 0)
    one_empty = ((ca == 0) | (cb == 0)) & ~both_empty
    out = np.where(union > 0, inter / np.maximum(union, 1.0), 0.0)
    out = np.where(one_empty, 0.0, out)
    return np.where(both_empty, 1.0, out)


def pair_distances(inter, cnt_a, cnt_b, label_range):
    """d[i][j] = 1 - mean over label_range of IoU(a_i == l, b_j == l); inter [Ka][Kb][L], cnt_a [Ka][L], cnt_b [Kb][L]"""
    lr = list(label_range)
    iou = _iou(inter[:, :, lr], cnt_a[:, None, lr], cnt_b[None, :, lr])
    return 1.0 - iou.sum(axis=-1) / len(lr)


def ged_from_counts(inter_sy, inter_ss, inter_yy, cnt_s, cnt_y, label_range):
    """2/(NM) sum d(s_i, y_j) - 1/N^2 sum d(s_i, s_j) - 1/M^2 sum d(y_i, y_j)   (utils.py:296-320)"""
    d_sy = pair_distances(inter_sy, cnt_s, cnt_y, label_range)
    d_ss = pair_distances(inter_ss, cnt_s, cnt_s, label_range)
    d_yy = pair_distances(inter_yy, cnt_y, cnt_y, label_range)
    n, m = d_sy.shape
    return float(2.0 / (n * m) * d_sy.sum() - d_ss.sum() / n ** 2 - d_yy.sum() / m ** 2)


def ncc_from_sums(sums, npix):
    """mean over the annotations of ncc(E_ss, E_sy[j]) = Pearson correlation of the two maps (utils.py:103-118,356-362);
    sums [M][5] = (sum a, sum a^2, sum v, sum v^2, sum a v)"""
    s = np.asarray(sums, np.float64).reshape(-1, 5)
    n = float(npix)
    ma, mv = s[:, 0] / n, s[:, 2] / n
    va = np.maximum(s[:, 1] / n - ma * ma, 0.0)
    vv = np.maximum(s[:, 3] / n - mv * mv, 0.0)
    with np.errstate(divide='ignore', invalid='ignore'):
        r = (s[:, 4] / n - ma * mv) / (np.sqrt(va) * np.sqrt(vv))      # nan for a constant map, like the reference
    return float(np.mean(r))


def dice_from_counts(inter, cnt_a, cnt_b):
    """per-label Dice of one prediction against one annotation (phiseg_model.py:596-606): inter, cnt_a, cnt_b [L]"""
    inter = np.asarray(inter, np.float64).reshape(-1)
    ca = np.asarray(cnt_a, np.float64).reshape(-1)
    cb = np.asarray(cnt_b, np.float64).reshape(-1)
    both_empty = (ca == 0) & (cb == 0)
    one_empty = ((ca == 0) | (cb == 0)) & ~both_empty
    d = np.where(ca + cb > 0, 2.0 * inter / np.maximum(ca + cb, 1.0), 0.0)
    d = np.where(one_empty, 0.0, d)
    return np.where(both_empty, 1.0, d)
