"""numpy restatement of the reference's array preparation (TEST INFRASTRUCTURE).

* :func:`crop_norm_cutout`, :func:`crop_triplets`  -- `btsbot/alert_utils.py:54-78`, `:81-107`
* :func:`triplet_tail`                             -- numeric tail of ``make_triplet``, `btsbot/alert_utils.py:159-193`
* :func:`to_model_layout`                          -- `btsbot/inference_example.py:62-64`, `train.py:139-155`, `val.py:92-94`
"""
import numpy as np


def crop_norm_cutout(cutout, crop_to_size):
    margin = (63 - crop_to_size) // 2
    cutout = cutout[margin:margin + crop_to_size, margin:margin + crop_to_size]
    cutout /= np.linalg.norm(cutout)          # in place on a view, as the reference does
    return cutout


def crop_triplets(triplets, crop_to_size):
    out = np.zeros((len(triplets), crop_to_size, crop_to_size, 3))
    for i in range(len(triplets)):
        for c in range(3):
            out[i, :, :, c] = crop_norm_cutout(triplets[i, :, :, c], crop_to_size)
    return out


def triplet_tail(stamps, normalize=True):
    """``stamps``: three 2-D float32 arrays (science, template, difference), NaN allowed, each <= 63x63.
    Returns ``(triplet[63,63,3] float64, drop)`` following alert_utils.py:147-193 (median check, nan_to_num,
    L2 normalise unless dropped, all-zero check, bottom/right pad with 1e-9 after normalisation)."""
    drop = False
    planes = []
    for data in stamps:
        with np.errstate(all="ignore"):
            median = np.nanmedian(data.flatten())
        if median == np.nan or median == -np.inf or median == np.inf:
            drop = True
        d = np.nan_to_num(data)
        if normalize and not drop:
            with np.errstate(all="ignore"):
                d = d / np.linalg.norm(d)
        if np.all(d.flatten() == 0):
            drop = True
        if d.shape != (63, 63):
            d = np.pad(d, [(0, 63 - d.shape[0]), (0, 63 - d.shape[1])], mode="constant", constant_values=1e-9)
        planes.append(d)
    trip = np.zeros((63, 63, 3))
    for c in range(3):
        trip[:, :, c] = planes[c]
    return trip, drop


def to_model_layout(triplets):
    """HWC float -> contiguous NCHW float32."""
    return np.ascontiguousarray(np.transpose(triplets.astype(np.float32), (0, 3, 1, 2)))
