"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's input pipeline:
data/batch_provider.py:43-67,124-272 and the OpenCV helpers it calls through utils.py:18-37.  Only tests/ may import it.

PARITY UNPINNED: OpenCV (cv2, a third-party dependency of the reference, version not pinned by the reference; absent
from this image) does the resampling.  The two functions below restate the published algorithm of OpenCV's
modules/imgproc/src/imgwarp.cpp (cv::warpAffine -> WarpAffineInvoker -> remapBilinear) and resize.cpp (cv::resize,
INTER_LINEAR, HResizeLinear / VResizeLinear) for floating-point images:
  * warpAffine inverts the matrix in double, evaluates source coordinates in fixed point (AB_BITS = 10, rounding with
    cvRound = round-half-even, + AB_SCALE / INTER_TAB_SIZE / 2), keeps 5 fractional bits (INTER_BITS) and blends the four
    neighbours with float table weights (1 - f) * (1 - g) ..., taps outside the image contributing the border value 0;
  * resize computes fx = (float)((dx + 0.5) * scale - 0.5) with scale = 1 / (dst / src), floors it, zeroes the fraction
    where the tap is clamped along x, clamps only the ROW indices along y, blends horizontally then vertically.
Whole images at a time (the CUDA kernel works per output pixel), in the image's own floating type like cv2 (float32 ->
float32 work type, float64 -> float64), weights in float32.
"""
import math

import numpy as np


# ---------------------------------------------------------------------------- OpenCV restatement
def get_rotation_matrix_2d(center, angle, scale):
    """cv2.getRotationMatrix2D: center is a Point2f; angle in degrees, positive = counter-clockwise"""
    cx, cy = float(np.float32(center[0])), float(np.float32(center[1]))
    a = angle * math.pi / 180.0
    alpha, beta = math.cos(a) * scale, math.sin(a) * scale
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]], np.float64)


def _invert_affine(M):
    m = [float(v) for v in np.asarray(M, np.float64).ravel()]
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2] = b1; m[5] = b2
    return m


def warp_affine_linear(img, M):
    """cv2.warpAffine(img, M, (cols, rows), flags=cv2.INTER_LINEAR) (borderMode constant, value 0); img [H,W] or [H,W,C]"""
    img = np.asarray(img)
    wt = np.float32 if img.dtype == np.float32 else np.float64
    src = img.astype(wt).reshape(img.shape[0], img.shape[1], -1)
    H, W, C = src.shape
    m = _invert_affine(M)
    xs = np.arange(W, dtype=np.float64)
    ys = np.arange(H, dtype=np.float64)
    adelta = np.rint(m[0] * xs * 1024.0).astype(np.int64)
    bdelta = np.rint(m[3] * xs * 1024.0).astype(np.int64)
    X0 = np.rint((m[1] * ys + m[2]) * 1024.0).astype(np.int64) + 16
    Y0 = np.rint((m[4] * ys + m[5]) * 1024.0).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    ix, iy = X >> 5, Y >> 5
    fx = (X & 31).astype(np.float32) * np.float32(1 / 32)
    fy = (Y & 31).astype(np.float32) * np.float32(1 / 32)
    wx = [np.float32(1) - fx, fx]
    wy = [np.float32(1) - fy, fy]
    out = None
    for k1 in range(2):
        for k2 in range(2):
            yy, xx = iy + k1, ix + k2
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            v = np.where(ok[..., None], src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], wt(0))
            term = v * (wy[k1] * wx[k2]).astype(wt)[..., None]
            out = term if out is None else out + term
    return out.reshape(img.shape).astype(img.dtype if img.dtype in (np.float32, np.float64) else wt)


def resize_linear(img, dsize_wh):
    """cv2.resize(img, (width, height), interpolation=cv2.INTER_LINEAR) for a floating-point image [H,W] or [H,W,C]"""
    img = np.asarray(img)
    wt = np.float32 if img.dtype == np.float32 else np.float64
    src = img.astype(wt).reshape(img.shape[0], img.shape[1], -1)
    sh, sw, C = src.shape
    dw, dh = int(dsize_wh[0]), int(dsize_wh[1])

    def taps(dst, ssz, along_x):
        scale = 1.0 / (float(dst) / float(ssz))
        f = ((np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = f - s.astype(np.float32)
        if along_x:
            lo = s < 0
            f[lo] = 0; s[lo] = 0
            hi = s >= ssz - 1
            f[hi] = 0; s[hi] = ssz - 1
            i0, i1 = s, np.minimum(s + 1, ssz - 1)
        else:
            i0, i1 = np.clip(s, 0, ssz - 1), np.clip(s + 1, 0, ssz - 1)
        return i0, i1, (np.float32(1) - f).astype(wt), f.astype(wt)

    c0, c1, a0, a1 = taps(dw, sw, True)
    r0, r1, b0, b1 = taps(dh, sh, False)
    hor = src[:, c0] * a0[None, :, None] + src[:, c1] * a1[None, :, None]          # [sh, dw, C]
    out = hor[r0] * b0[:, None, None] + hor[r1] * b1[:, None, None]
    shape = (dh, dw) + img.shape[2:]
    return out.reshape(shape)


# ---------------------------------------------------------------------------- utils.py:18-37,86-91
def convert_to_onehot(lblmap, nlabels):
    out = np.zeros((lblmap.shape[0], lblmap.shape[1], nlabels))
    for ii in range(nlabels):
        out[:, :, ii] = (lblmap == ii).astype(np.uint8)
    return out


def rotate_image(img, angle):
    rows, cols = img.shape[:2]
    return warp_affine_linear(img, get_rotation_matrix_2d((cols / 2, rows / 2), angle, 1))


def rotate_image_as_onehot(img, angle, nlabels):
    return np.argmax(rotate_image(convert_to_onehot(img, nlabels), angle), axis=-1)


def resize_image(im, size):
    return resize_linear(im, (size[1], size[0]))


def resize_image_as_onehot(im, size, nlabels):
    return np.argmax(resize_image(convert_to_onehot(im, nlabels), size), axis=-1)


# ---------------------------------------------------------------------------- data/batch_provider.py
class BatchProvider:
    """next_batch of data/batch_provider.py:43-67 with _select_random_label (:124-130) and _augmentation_function
    (:133-272, rotation / crop-scale / flips; nlabels <= 4), drawing from the global np.random like the reference."""

    def __init__(self, X, y, indices, add_dummy_dimension=False, **kwargs):
        self.X, self.y = X, y
        self.indices = indices
        self.unused_indices = indices.copy()
        self.add_dummy_dimension = add_dummy_dimension
        self.num_labels_per_subject = kwargs.get('num_labels_per_subject', 1)
        if self.num_labels_per_subject > 1:
            self.annotator_range = kwargs.get('annotator_range', range(self.num_labels_per_subject))
        self.do_augmentations = kwargs.get('do_augmentations', False)
        self.augmentation_options = kwargs.get('augmentation_options', None)

    def next_batch(self, batch_size):
        if len(self.unused_indices) < batch_size:
            self.unused_indices = self.indices
        batch_indices = np.random.choice(self.unused_indices, batch_size, replace=False)
        self.unused_indices = np.setdiff1d(self.unused_indices, batch_indices)
        batch_indices = np.sort(batch_indices)
        X_batch = self.X[batch_indices, ...]
        y_batch = self.y[batch_indices, ...]
        if self.num_labels_per_subject > 1:
            y_batch = np.asarray([y_batch[ii, ..., np.random.choice(self.annotator_range)] for ii in range(y_batch.shape[0])])
        if self.do_augmentations:
            X_batch, y_batch = self._augment(X_batch, y_batch)
        if self.add_dummy_dimension:
            X_batch = np.expand_dims(X_batch, axis=-1)
        return X_batch, y_batch

    def _augment(self, images, labels):
        opt = self.augmentation_options
        get = lambda name, default: opt[name] if name in opt else default
        do_rotations, do_scaleaug = get('do_rotations', False), get('do_scaleaug', False)
        do_fliplr, do_flipud = get('do_fliplr', False), get('do_flipud', False)
        nth = get('augment_every_nth', 2)
        nlabels = get('nlabels', None)
        new_images, new_labels = [], []
        for ii in range(images.shape[0]):
            img = np.squeeze(images[ii, ...])
            lbl = np.squeeze(labels[ii, ...])
            if np.random.randint(nth) == 0:
                if do_rotations:
                    angles = get('rot_degrees', 10.0)
                    random_angle = np.random.uniform(-angles, angles)
                    img = rotate_image(img, random_angle)
                    lbl = rotate_image_as_onehot(lbl, random_angle, nlabels=nlabels)
                if do_scaleaug:
                    offset = get('offset', 30)
                    n_x, n_y = img.shape
                    r_y = np.random.randint(n_y - offset, n_y + 1)          # random_integers(lo, hi) = randint(lo, hi + 1)
                    p_x = np.random.randint(0, n_x - r_y + 1)
                    p_y = np.random.randint(0, n_y - r_y + 1)
                    img = resize_image(img[p_y:(p_y + r_y), p_x:(p_x + r_y)], (n_x, n_y))
                    lbl = resize_image_as_onehot(lbl[p_y:(p_y + r_y), p_x:(p_x + r_y)], (n_x, n_y), nlabels=nlabels)
            if do_fliplr and np.random.randint(max(2, nth)) == 0:
                img, lbl = np.fliplr(img), np.fliplr(lbl)
            if do_flipud and np.random.randint(max(2, nth)) == 0:
                img, lbl = np.flipud(img), np.flipud(lbl)
            new_images.append(img[...])
            new_labels.append(lbl[...])
        return np.asarray(new_images), np.asarray(new_labels)
