"""Seeded synthetic batches of the shape the reference's batch provider yields (data/batch_provider.py:43-67):
x float32 [B,H,W,1] in [-0.5, 0.5] (data/lidc_data_loader.py:92 stores image - 0.5), s uint8 [B,H,W] label masks.
Used by bench.py and the tools.  The CPU checker under tests/ carries its own identical generator so that neither side
imports the other (a CPU test checks that the two agree bit for bit)."""
import numpy as np


def synthetic_batch(B, H=128, W=128, nlabels=2, seed=1234):
    """Smooth random field images and random ellipse 'lesions' (about one image in five stays empty)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, H, W)).astype(np.float32)
    k = np.exp(-0.5 * (np.arange(-6, 7) / 2.5) ** 2)
    k /= k.sum()
    for ax in (1, 2):
        x = np.apply_along_axis(lambda v: np.convolve(v, k, mode='same'), ax, x)
    x = np.clip(x / (3 * x.std() + 1e-8), -0.5, 0.5).astype(np.float32)[..., None]
    yy, xx = np.mgrid[0:H, 0:W]
    s = np.zeros((B, H, W), np.uint8)
    for b in range(B):
        if rng.random() < 0.2:
            continue
        for lab in range(1, nlabels):
            cy, cx = rng.uniform(0.3 * H, 0.7 * H), rng.uniform(0.3 * W, 0.7 * W)
            ry, rx = rng.uniform(0.04 * H, 0.16 * H) / lab, rng.uniform(0.04 * W, 0.16 * W) / lab
            s[b][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = lab
    return x, s


def synthetic_eps(shapes, seed=1234):
    """One N(0,1) array per latent level (stands for the unseeded tf.random_normal of posteriors.py:108)."""
    rng = np.random.default_rng(seed + 77)
    return [rng.standard_normal(sh).astype(np.float32) for sh in shapes]


class _Split:
    """images [N,H,W] float32, labels [N,H,W,A] uint8 with the next_batch interface of data/batch_provider.py:43-67
    (a random annotator per image, data/batch_provider.py:103-110)"""

    def __init__(self, images, labels, seed):
        self.images, self.labels = images, labels
        self._rng = np.random.default_rng(seed)

    def next_batch(self, batch_size):
        idx = self._rng.choice(self.images.shape[0], size=batch_size, replace=self.images.shape[0] < batch_size)
        ann = self._rng.integers(0, self.labels.shape[-1], size=batch_size)
        x = self.images[idx][..., None].astype(np.float32)
        s = np.stack([self.labels[i, :, :, a] for i, a in zip(idx, ann)])
        return x, s.astype(np.uint8)


class SyntheticLIDC:
    """Stand-in for data/lidc_data.py: .train / .validation splits of smooth random images with `annotators` jittered
    lesion masks each (the real LIDC-IDRI crops are not available offline)."""

    def __init__(self, num_train=64, num_val=8, size=128, nlabels=2, annotators=4, seed=0):
        def make(n, sd):
            x, _ = synthetic_batch(n, size, size, nlabels, seed=sd)
            labs = np.stack([synthetic_batch(n, size, size, nlabels, seed=sd)[1]] +
                            [np.roll(synthetic_batch(n, size, size, nlabels, seed=sd)[1], shift=a, axis=1 + a % 2)
                             for a in range(1, annotators)], axis=-1)
            return x[..., 0], labs.astype(np.uint8)
        self.train = _Split(*make(num_train, seed), seed=seed + 1)
        self.validation = _Split(*make(num_val, seed + 100), seed=seed + 2)
