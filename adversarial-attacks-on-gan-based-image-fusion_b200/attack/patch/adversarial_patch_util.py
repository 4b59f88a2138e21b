"""The helpers `attack/patch/adversarial_patch.py` imports from the (missing) `adversarial_patch_util` module
(call sites: adversarial_patch.py:40-42,67,203-207,217-219,233-235).  Restated from those call sites: square patches only
(the circle variant needs scipy rotations and is never used by the configs of BASELINE.json)."""
import numpy as np


def init_patch_square(image_size, patch_size):
    """-> (patch ndarray (1,3,d,d) in [0,1), patch_shape); d = floor(sqrt(image_size^2 * patch_size))  (adversarial_patch.py:216-219)"""
    d = int((image_size * image_size * patch_size) ** 0.5)
    patch = np.random.rand(1, 3, d, d)
    return patch, patch.shape


def square_transform(patch, data_shape, patch_shape, image_size, centre=True, rng=None):
    """place the patch on a zero canvas of `data_shape`; returns (canvas, mask)  (adversarial_patch.py:41-42)"""
    x = np.zeros(data_shape, dtype=np.float32)
    d = patch_shape[-1]
    for i in range(x.shape[0]):
        if centre:
            r = c = (image_size - d) // 2
        else:
            rng = rng or np.random
            r, c = rng.randint(0, image_size - d + 1), rng.randint(0, image_size - d + 1)
        x[i, :, r:r + d, c:c + d] = patch[0]
    mask = (x != 0).astype(np.float32)
    mask[:, :, :, :] = 0
    for i in range(x.shape[0]):
        r = c = (image_size - d) // 2 if centre else 0
        mask[i, :, r:r + d, c:c + d] = 1.0
    return x, mask


def submatrix(arr):
    """crop a 2-D array to the bounding box of its non-zero entries (adversarial_patch.py:62-67)"""
    x, y = np.nonzero(arr)
    if len(x) == 0:
        return arr
    return arr[x.min():x.max() + 1, y.min():y.max() + 1]
