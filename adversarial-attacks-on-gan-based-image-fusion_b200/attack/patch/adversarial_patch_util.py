"""The helpers `code/attack/patch/adversarial_patch.py` star-imports from the module `attack.patch.adversarial_patch_util`, which
the reference does not ship.  Its call sites (adversarial_patch.py:40-42 `circle_transform`/`square_transform`,
:67 `submatrix`, :203-207 and :217-219 `init_patch_circle`/`init_patch_square`, :233-235) match the helper module of the public
adversarial-patch training script the reference's `train`/`attack` loop is modelled on; what each helper must do follows from
those call sites and is restated here:

  init_patch_square(image_size, patch_size)  -> (patch (1,3,d,d) ~ U[0,1), shape),  d = floor(sqrt(image_size^2 * patch_size))
  init_patch_circle(image_size, patch_size)  -> (patch (1,3,2r,2r), shape): a disc of radius r = floor(sqrt(area/pi)), one random
                                                value per colour plane, zero outside the disc
  square_transform(patch, data_shape, patch_shape, image_size) -> (canvas, mask): per image a random multiple of 90 degrees and a
                                                random position that keeps the patch inside the image; mask = (canvas != 0)
  circle_transform(...)                      -> (canvas, mask, patch_shape): same with a random rotation angle
  submatrix(arr)                             -> arr cropped to the bounding box of its non-zero entries (carries the patch from one
                                                image to the next, adversarial_patch.py:62-69)

Host-side numpy by design: this is the once-per-image placement glue around the iterated loop, not the loop.
`rng` (a numpy Generator / RandomState, default the global numpy state as upstream) and `centre=True` (BASELINE config 4 fixes the
patch at the image centre) are additions; the defaults keep the reference behaviour.
"""
import math

import numpy as np


def _rng(rng):
    return np.random if rng is None else rng


def _randint(rng, hi):
    r = _rng(rng)
    return int(r.integers(hi)) if hasattr(r, "integers") else int(r.randint(hi))


def _uniform(rng, shape=None):
    r = _rng(rng)
    return r.random(shape) if hasattr(r, "integers") else r.random_sample(shape)


def init_patch_square(image_size, patch_size, rng=None):
    d = int((image_size * image_size * patch_size) ** 0.5)
    patch = _uniform(rng, (1, 3, d, d))
    return patch, patch.shape


def init_patch_circle(image_size, patch_size, rng=None):
    area = int(image_size * image_size * patch_size)
    radius = int(math.sqrt(area / math.pi))
    yy, xx = np.ogrid[-radius:radius, -radius:radius]
    disc = (xx * xx + yy * yy) <= radius * radius
    patch = np.zeros((1, 3, 2 * radius, 2 * radius))
    for c in range(3):
        patch[0, c][disc] = float(_uniform(rng))
    return patch, patch.shape


def _place(patch, data_shape, patch_shape, image_size, rotate, centre, rng):
    canvas = np.zeros(data_shape, dtype=np.float64)
    d = patch_shape[-1]
    assert d <= image_size, "patch larger than the image"
    placed = []
    for i in range(canvas.shape[0]):
        p = patch[i if i < patch.shape[0] else 0]
        p = np.stack([rotate(p[c]) for c in range(p.shape[0])]) if rotate is not None else p
        if centre:
            r = c = (image_size - d) // 2
        else:
            r, c = _randint(rng, image_size - d + 1), _randint(rng, image_size - d + 1)
        canvas[i, :, r:r + d, c:c + d] = p
        placed.append((r, c))
    mask = (canvas != 0).astype(np.float64)        # the patch IS the mask: whatever was pasted (non-zero) is optimised
    return canvas, mask, placed


def square_transform(patch, data_shape, patch_shape, image_size, centre=False, rng=None):
    k = None if centre else _randint(rng, 4)
    rot = (lambda a: np.rot90(a, k)) if k else None
    canvas, mask, _ = _place(np.asarray(patch), tuple(data_shape), tuple(patch_shape), image_size, rot, centre, rng)
    return canvas, mask


def _rotate_nearest(a, angle_deg):
    """rotation about the centre, same output shape, zero fill, bilinear (what a reshape=False image rotation does)"""
    h, w = a.shape
    t = math.radians(angle_deg)
    cy, cx = (h - 1) / 2.0, (w - 1) / 2.0
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    ys = cy + (yy - cy) * math.cos(t) - (xx - cx) * math.sin(t)
    xs = cx + (yy - cy) * math.sin(t) + (xx - cx) * math.cos(t)
    y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
    fy, fx = ys - y0, xs - x0
    out = np.zeros_like(a, dtype=np.float64)
    for dy, wy in ((0, 1 - fy), (1, fy)):
        for dx, wx in ((0, 1 - fx), (1, fx)):
            yi, xi = y0 + dy, x0 + dx
            ok = (yi >= 0) & (yi < h) & (xi >= 0) & (xi < w)
            out += np.where(ok, a[np.clip(yi, 0, h - 1), np.clip(xi, 0, w - 1)], 0.0) * wy * wx
    return out


def circle_transform(patch, data_shape, patch_shape, image_size, centre=False, rng=None):
    ang = 0 if centre else _randint(rng, 360)
    rot = (lambda a: _rotate_nearest(a, ang)) if ang else None
    canvas, mask, _ = _place(np.asarray(patch), tuple(data_shape), tuple(patch_shape), image_size, rot, centre, rng)
    return canvas, mask, tuple(patch_shape)


def submatrix(arr):
    rows, cols = np.nonzero(arr)
    if rows.size == 0:
        return arr
    return arr[rows.min():rows.max() + 1, cols.min():cols.max() + 1]
