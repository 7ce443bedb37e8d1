"""Device-resident counterpart of data/batch_provider.py::BatchProvider (same constructor keywords, same next_batch /
iterate_batches interface, same np.random call order) - SURVEY.md section 8f N2.

The reference gathers a batch from HDF5 on the host, picks a random annotator per image and augments image by image
with OpenCV (batch_provider.py:43-67,124-272); at several thousand images per second per GPU that loop cannot keep up.
Here the data set is uploaded ONCE, the host only draws the random numbers - with the reference's calls in the reference's
order, so np.random.seed(k) selects the same images, annotators and augmentation parameters - and one kernel launch
(phs_augment_batch) gathers, augments and converts the batch on the device.  next_batch() returns host arrays like the
reference; next_batch_device() returns CUDA tensors that phiseg.training_step consumes without a host round trip.

Options of the reference that its LIDC pipeline never switches on (resize_to, rescale_range, rescale_rgb,
do_elasticaug, more than 4 labels) raise ValueError instead of silently doing something else.  The reference's
normalise_images call discards its result (batch_provider.py:118), so no normalisation happens here either.
"""
import ctypes
import math

import numpy as np
import torch

from .. import lib as L

PARAM_DTYPE = np.dtype([('src', '<i4'), ('annot', '<i4'), ('flags', '<i4'), ('crop', '<i4'), ('px', '<i4'), ('py', '<i4'),
                        ('minv', '<f8', (6,))])
assert PARAM_DTYPE.itemsize == ctypes.sizeof(L.phs_aug_params)


def inverse_rotation_matrix(rows, cols, angle_deg):
    """The 2x3 matrix cv2.warpAffine maps destination to source pixels with, for utils.rotate_image (utils.py:18-22):
    cv2.getRotationMatrix2D((cols / 2, rows / 2), angle, 1), inverted the way warpAffine inverts it."""
    cx, cy = np.float32(cols / 2), np.float32(rows / 2)          # the centre is a Point2f
    a = angle_deg * math.pi / 180.0
    alpha, beta = math.cos(a), math.sin(a)
    m = [alpha, beta, (1 - alpha) * float(cx) - beta * float(cy), -beta, alpha, beta * float(cx) + (1 - alpha) * float(cy)]
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0], m[1], m[3], m[4] = a11, m[1] * -d, m[3] * -d, a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


class BatchProvider:
    """X [N,H,W] float32 / float64 images, y [N,H,W,A] uint8 annotation masks (or [N,H,W]); indices: the rows this
    provider may draw (data/batch_provider.py:25-41)."""

    def __init__(self, X, y, indices, add_dummy_dimension=False, device=None, **kwargs):
        self.lib = L.load()
        self.device = torch.device(device if device is not None else 'cuda')
        X = np.asarray(X)
        y = np.asarray(y)
        if X.ndim != 3:
            raise ValueError('X must be [N,H,W] (2D single-channel images), got %s' % (X.shape,))
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float32)
        self.num_labels_per_subject = kwargs.get('num_labels_per_subject', 1)
        if y.ndim == 3:
            y = y[..., None]
        if y.shape[:3] != X.shape or y.ndim != 4:
            raise ValueError('y must be [N,H,W] or [N,H,W,annotators] matching X, got %s' % (y.shape,))
        self.annotator_range = list(kwargs.get('annotator_range', range(self.num_labels_per_subject)))
        if max(self.annotator_range) >= y.shape[-1]:
            raise ValueError('annotator_range exceeds the %d annotations per image' % y.shape[-1])
        self.indices = np.asarray(indices).copy()
        self.unused_indices = self.indices.copy()
        self.add_dummy_dimension = add_dummy_dimension
        for unsupported in ('resize_to', 'rescale_range', 'rescale_rgb'):
            if kwargs.get(unsupported):
                raise ValueError('%s is not part of the LIDC pipeline and is not implemented' % unsupported)
        self.do_augmentations = kwargs.get('do_augmentations', False)
        self.augmentation_options = kwargs.get('augmentation_options', None) or {}
        if self.do_augmentations:
            opt = self.augmentation_options
            if opt.get('do_elasticaug', False):
                raise ValueError('do_elasticaug is not implemented')
            if opt.get('do_rotations', False) or opt.get('do_scaleaug', False):
                if not opt.get('nlabels', None):
                    raise AssertionError("When doing augmentations with rotations, scaling, or elastic transformations "
                                         "the parameter 'nlabels' must be provided.")      # batch_provider.py:166-170
                if opt['nlabels'] > 4:
                    raise ValueError('label interpolation is implemented for nlabels <= 4 (one-hot branch)')
        self.nlabels = int(self.augmentation_options.get('nlabels', 0) or (int(y.max()) + 1 if y.size else 1))
        if y.size and int(y.max()) >= max(self.nlabels, 1) and self.do_augmentations:
            raise ValueError('labels must lie in [0, %d)' % self.nlabels)
        self.N, self.H, self.W = X.shape
        self.A = y.shape[-1]
        # the resident data set
        self.X = torch.as_tensor(np.ascontiguousarray(X)).to(self.device)
        self.y = torch.as_tensor(np.ascontiguousarray(y.astype(np.uint8))).to(self.device)
        self._slots = {}
        self.launches = 0

    # ---- random parameters: the reference's np.random calls in the reference's order ---------------------------
    def _draw_indices(self, batch_size):
        """batch_provider.py:49-56: sampling without replacement across batches"""
        if len(self.unused_indices) < batch_size:
            self.unused_indices = self.indices
        batch_indices = np.random.choice(self.unused_indices, batch_size, replace=False)
        self.unused_indices = np.setdiff1d(self.unused_indices, batch_indices)
        return np.sort(batch_indices)                                     # 'HDF5 requires indices to be in increasing order'

    def _draw_params(self, batch_indices):
        B = len(batch_indices)
        p = np.zeros(B, dtype=PARAM_DTYPE)
        p['src'] = batch_indices
        p['crop'] = self.W
        if self.num_labels_per_subject > 1:                               # _select_random_label, :124-130
            for ii in range(B):
                p['annot'][ii] = np.random.choice(self.annotator_range)
        if not self.do_augmentations:
            return p
        opt = self.augmentation_options
        get = lambda name, default: opt[name] if name in opt else default
        do_rot, do_scale = get('do_rotations', False), get('do_scaleaug', False)
        do_fliplr, do_flipud = get('do_fliplr', False), get('do_flipud', False)
        nth = get('augment_every_nth', 2)
        n_x, n_y = self.H, self.W                                         # img.shape of one image
        for ii in range(B):                                               # :179-247
            flags = 0
            if np.random.randint(nth) == 0:
                if do_rot:
                    angles = get('rot_degrees', 10.0)
                    angle = np.random.uniform(-angles, angles)
                    p['minv'][ii] = inverse_rotation_matrix(self.H, self.W, angle)
                    flags |= L.AUG_ROTATE
                if do_scale:
                    offset = get('offset', 30)
                    r_y = np.random.randint(n_y - offset, n_y + 1)        # np.random.random_integers(lo, hi): hi inclusive
                    p_x = np.random.randint(0, n_x - r_y + 1)
                    p_y = np.random.randint(0, n_y - r_y + 1)
                    p['crop'][ii], p['px'][ii], p['py'][ii] = r_y, p_x, p_y
                    flags |= L.AUG_SCALE
            if do_fliplr and np.random.randint(max(2, nth)) == 0:
                flags |= L.AUG_FLIPLR
            if do_flipud and np.random.randint(max(2, nth)) == 0:
                flags |= L.AUG_FLIPUD
            p['flags'][ii] = flags
        return p

    # ---- device work --------------------------------------------------------------------------------------------
    def _slot(self, B):
        s = self._slots.get(B)
        if s is None:
            s = {'k': 0, 'h': [torch.zeros(B * PARAM_DTYPE.itemsize, dtype=torch.uint8).pin_memory() for _ in range(2)],
                 'd': [torch.empty(B * PARAM_DTYPE.itemsize, dtype=torch.uint8, device=self.device) for _ in range(2)],
                 'x': [torch.empty((B, self.H, self.W, 1), dtype=torch.float32, device=self.device) for _ in range(2)],
                 's': [torch.empty((B, self.H, self.W), dtype=torch.uint8, device=self.device) for _ in range(2)],
                 'ev': [torch.cuda.Event() for _ in range(2)]}
            self._slots[B] = s
        return s

    def _launch(self, params):
        """params -> (x [B,H,W,1] float32, s [B,H,W] uint8) CUDA tensors; two rotating output slots, so the tensors of
        one call stay valid until the call after the next."""
        B = len(params)
        sl = self._slot(B)
        j = sl['k'] & 1
        sl['k'] += 1
        sl['ev'][j].synchronize()                       # the slot's previous parameter upload has left pinned memory
        sl['h'][j].numpy()[:] = params.view(np.uint8)
        sl['d'][j].copy_(sl['h'][j], non_blocking=True)
        sl['ev'][j].record()
        st = torch.cuda.current_stream().cuda_stream
        L.check(self.lib.phs_augment_batch(self.X.data_ptr(), L.PHS_F64 if self.X.dtype == torch.float64 else L.PHS_F32,
                                           self.y.data_ptr(), self.H, self.W, self.A, max(self.nlabels, 1),
                                           sl['d'][j].data_ptr(), B, sl['x'][j].data_ptr(), sl['s'][j].data_ptr(), st),
                'phs_augment_batch')
        self.launches += 1
        return sl['x'][j], sl['s'][j]

    def next_batch_device(self, batch_size):
        """One random batch as CUDA tensors: x [B,H,W,1] float32, s [B,H,W] uint8."""
        return self._launch(self._draw_params(self._draw_indices(batch_size)))

    def _to_host(self, x, s):
        x = x.cpu().numpy()
        return (x if self.add_dummy_dimension else x[..., 0]), s.cpu().numpy()

    def next_batch(self, batch_size):
        """batch_provider.py:43-67 (host arrays, like the reference)."""
        return self._to_host(*self.next_batch_device(batch_size))

    def iterate_batches(self, batch_size, shuffle=True):
        """batch_provider.py:69-99: one pass over the indices (the last batch may be short)."""
        if shuffle:
            np.random.shuffle(self.indices)
        for b_i in range(0, self.indices.shape[0], batch_size):
            batch_indices = np.sort(self.indices[b_i:b_i + batch_size])
            yield self._to_host(*self._launch(self._draw_params(batch_indices)))


class lidc_data:
    """data/lidc_data.py:8-52 over in-memory arrays: data = {'train' | 'val' | 'test': {'images': [N,H,W], 'labels':
    [N,H,W,A]}} (what lidc_data_loader.load_and_maybe_process_data returns; h5py is not needed for arrays or .npz)."""

    def __init__(self, exp_config, data, device=None):
        self.data = data
        if not hasattr(exp_config, 'annotator_range'):
            exp_config.annotator_range = range(exp_config.num_labels_per_subject)
        common = dict(add_dummy_dimension=True, num_labels_per_subject=exp_config.num_labels_per_subject,
                      annotator_range=exp_config.annotator_range, device=device)
        idx = {tt: np.arange(np.shape(data[tt]['images'])[0]) for tt in data}
        self.train = BatchProvider(data['train']['images'], data['train']['labels'], idx['train'], do_augmentations=True,
                                   augmentation_options=exp_config.augmentation_options, **common)
        self.validation = BatchProvider(data['val']['images'], data['val']['labels'], idx['val'], **common)
        self.validation.images, self.validation.labels = data['val']['images'], data['val']['labels']
        if 'test' in data:
            self.test = BatchProvider(data['test']['images'], data['test']['labels'], idx['test'], **common)
            self.test.images, self.test.labels = data['test']['images'], data['test']['labels']
