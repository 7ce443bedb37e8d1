"""Input pipeline (data/batch_provider.py:43-67,124-272, utils.py:18-37).

CPU: hand-derived known answers for the oracle's OpenCV restatement (cv2 is not installed: parity unpinned, see
oracle/input_pipeline.py), and the product's random-parameter drawing against the oracle's np.random consumption.
GPU (-m gpu): same np.random seed -> the device provider's batches equal the oracle's numpy pipeline bit for bit
(float32 and float64 data sets, 2 and 4 labels, rotations + crop-scaling + flips), and training_step on device tensors."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def ip():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    try:
        return importlib.import_module('input_pipeline')
    finally:
        sys.path.pop(0)


def _dataset(N=24, H=32, W=32, A=4, nl=2, dtype=np.float64, seed=0):
    rng = np.random.default_rng(seed)
    X = (rng.random((N, H, W)) - 0.5).astype(dtype)
    yy, xx = np.mgrid[0:H, 0:W]
    y = np.zeros((N, H, W, A), np.uint8)
    for n in range(N):
        for a in range(A):
            for lab in range(1, nl):
                cy, cx, r = rng.uniform(8, H - 8), rng.uniform(8, W - 8), rng.uniform(2, 9)
                y[n, :, :, a][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = lab
    return X, y


OPTS = {'do_rotations': True, 'do_scaleaug': True, 'do_fliplr': True, 'do_flipud': True, 'offset': 10, 'rot_degrees': 15.0}


def test_opencv_restatement_known_answers(ip):
    rng = np.random.default_rng(1)
    img = rng.random((8, 8))
    # angle 0: the fixed-point coordinates land on the pixels themselves
    assert np.array_equal(ip.rotate_image(img, 0.0), img)
    # 90 degrees about (cols/2, rows/2) = (4, 4): dst[y, x] = src[x, 8 - y]; row 0 looks at column 8 = border value 0
    r = ip.rotate_image(img, 90.0)
    want = np.zeros_like(img)
    for y in range(1, 8):
        for x in range(8):
            want[y, x] = img[x, 8 - y]
    assert np.array_equal(r, want)
    # half-pixel shift: M = [[1, 0, .5], [0, 1, 0]] -> dst[x] = (src[x-1] + src[x]) / 2 with src[-1] = 0
    s = ip.warp_affine_linear(img, np.array([[1.0, 0, 0.5], [0, 1.0, 0]]))
    assert np.array_equal(s[:, 1:], 0.5 * img[:, :-1] + 0.5 * img[:, 1:]) and np.array_equal(s[:, 0], 0.5 * img[:, 0])
    # cv2.resize INTER_LINEAR of [0, 1] to four samples: the classic (0, .25, .75, 1); rows replicate
    z = ip.resize_linear(np.array([[0.0, 1.0]]), (4, 2))
    assert z.shape == (2, 4) and np.array_equal(z, np.array([[0, 0.25, 0.75, 1.0]] * 2))
    assert np.array_equal(ip.resize_linear(img, (8, 8)), img)
    # float32 images are processed in float32
    assert ip.rotate_image(img.astype(np.float32), 7.0).dtype == np.float32
    # labels: rotating a mask by 0 keeps it; a crop stretched back keeps the label set
    lbl = (img > 0.5).astype(np.uint8)
    assert np.array_equal(ip.rotate_image_as_onehot(lbl, 0.0, 2), lbl)
    assert set(np.unique(ip.resize_image_as_onehot(lbl[1:7, 1:7], (8, 8), 2))) <= {0, 1}


def test_parameter_drawing_follows_the_reference_rng_order(ip, pkg):
    """The device provider draws its random numbers with the calls and in the order of data/batch_provider.py: after
    np.random.seed(k) both sides have consumed the same stream and picked the same images / annotators."""
    bp = importlib.import_module('phiseg_code_b200.data.batch_provider')
    X, y = _dataset(N=10, H=16, W=16)
    opts = dict(OPTS, nlabels=2, offset=4)
    kw = dict(add_dummy_dimension=True, do_augmentations=True, augmentation_options=opts, num_labels_per_subject=4,
              annotator_range=range(4))
    dev = bp.BatchProvider(X, y, np.arange(10), device='cpu', **kw)
    ref = ip.BatchProvider(X, y, np.arange(10), **kw)
    np.random.seed(5)
    drawn = [dev._draw_params(dev._draw_indices(4)) for _ in range(4)]       # 4 x 4 of 10 indices: the pool refills once
    probe_dev = np.random.random()
    np.random.seed(5)
    for _ in range(4):
        ref.next_batch(4)
    assert np.random.random() == probe_dev
    # without augmentation the batch is a pure gather: check it against the parameters
    plain = bp.BatchProvider(X, y, np.arange(10), device='cpu', num_labels_per_subject=4, annotator_range=range(4))
    ref2 = ip.BatchProvider(X, y, np.arange(10), num_labels_per_subject=4, annotator_range=range(4))
    np.random.seed(9)
    p = plain._draw_params(plain._draw_indices(6))
    np.random.seed(9)
    xr, yr = ref2.next_batch(6)
    assert np.array_equal(xr, X[p['src']]) and np.array_equal(yr, np.stack([y[i, :, :, a] for i, a in zip(p['src'], p['annot'])]))
    assert all(d['flags'].max() <= 15 and (d['crop'] <= 16).all() and (d['crop'] >= 12).all() for d in drawn)
    with pytest.raises(ValueError):
        bp.BatchProvider(X, y, np.arange(10), device='cpu', do_augmentations=True, augmentation_options={'do_elasticaug': True})
    with pytest.raises(AssertionError):
        bp.BatchProvider(X, y, np.arange(10), device='cpu', do_augmentations=True, augmentation_options={'do_rotations': True})


@pytest.mark.gpu
@pytest.mark.parametrize('dtype,nl', [(np.float64, 2), (np.float32, 2), (np.float64, 4)])
def test_device_batches_equal_the_numpy_pipeline(ip, pkg, dtype, nl):
    bp = importlib.import_module('phiseg_code_b200.data.batch_provider')
    X, y = _dataset(N=24, H=32, W=32, nl=nl, dtype=dtype, seed=nl)
    opts = dict(OPTS, nlabels=nl)
    kw = dict(add_dummy_dimension=True, do_augmentations=True, augmentation_options=opts, num_labels_per_subject=4,
              annotator_range=range(4))
    dev = bp.BatchProvider(X, y, np.arange(24), **kw)
    ref = ip.BatchProvider(X, y, np.arange(24), **kw)
    np.random.seed(123)
    got = [dev.next_batch(8) for _ in range(5)]
    np.random.seed(123)
    want = [ref.next_batch(8) for _ in range(5)]
    augmented = 0
    for (xg, sg), (xw, sw) in zip(got, want):
        assert xg.shape == (8, 32, 32, 1) and xg.dtype == np.float32 and sg.shape == (8, 32, 32) and sg.dtype == np.uint8
        assert np.array_equal(xg, xw.astype(np.float32)), np.abs(xg - xw.astype(np.float32)).max()
        assert np.array_equal(sg, sw)
        augmented += sum(not any(np.array_equal(xw[i, ..., 0], X[n]) for n in range(24)) for i in range(8))
    assert augmented >= 8, 'hardly any image was augmented: the comparison would be vacuous'
    assert dev.launches == 5                    # one kernel launch per batch
    # a validation-style provider: no augmentation, every annotator plane reachable, iterate_batches covers all rows once
    val = bp.BatchProvider(X, y, np.arange(24), add_dummy_dimension=True, num_labels_per_subject=4, annotator_range=range(4))
    np.random.seed(1)
    rows = sum(xb.shape[0] for xb, _ in val.iterate_batches(10, shuffle=False))
    assert rows == 24


@pytest.mark.gpu
def test_training_step_consumes_device_batches(pkg):
    """phiseg.training_step on the CUDA tensors of next_batch_device == on the same batch as host arrays"""
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    D = importlib.import_module('phiseg_code_b200.data')
    exp = ex.load_experiment(ex.experiment_path('phiseg_7_5'))
    exp.image_size = (64, 64, 1)
    X, y = _dataset(N=16, H=64, W=64, nl=2, dtype=np.float64, seed=4)
    data = D.lidc_data(exp, {'train': {'images': X, 'labels': y}, 'val': {'images': X[:4], 'labels': y[:4]}})
    np.random.seed(3)
    xd, sd = data.train.next_batch_device(4)
    xh, sh = xd.cpu().numpy(), sd.cpu().numpy()
    losses = []
    for batch in ((xd, sd), (xh, sh)):
        model = pm.phiseg(exp, mode='fast', use_cuda_graph=False, seed=11)
        model._gen.manual_seed(5)
        losses.append(model.training_step(batch[0], batch[1], lr=1e-3))
    # same inputs either way; what remains is the run-to-run spread of the step (test_fast_mode_reproducible: <= 1.8e-7)
    assert np.isfinite(losses[0]) and abs(losses[0] - losses[1]) <= 1e-6 * abs(losses[0]), losses
