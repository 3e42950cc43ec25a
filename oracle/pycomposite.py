"""numpy restatement of the reference's layer flatten + background composite (TEST INFRASTRUCTURE ONLY).

Pinned against the reference's own functions executed in this container (tools/make_golden.py --para-gen ->
tests/golden/para_gen_cases.npz; tests/test_oracle_golden.py).
Follows para_gen.py:136-175 (flatten: later segments overwrite where their warped mask is non-zero, by
`x*msk_bg + y*msk_ob`) and para_gen.py:50-61 (add_bg: background where the mask equals 0), line by line.
"""
import numpy as np


def flatten(flows, rgbs, masks):
    flow_im, rgb2_im, msk2_im = flows[0].copy(), rgbs[0].copy(), masks[0].copy()
    for i in range(1, len(flows)):
        msk_ob = masks[i] != 0
        msk_bg = masks[i] == 0
        flow_im = flow_im * msk_bg[..., None] + flows[i] * msk_ob[..., None]
        rgb2_im = rgb2_im * msk_bg[..., None] + rgbs[i] * msk_ob[..., None]
        msk2_im = msk2_im * msk_bg + masks[i] * msk_ob
    return flow_im.astype(np.float32), rgb2_im.astype(np.uint8), msk2_im.astype(np.uint8)


def add_bg(im, mk, bgim, bgval=0):
    out = im.copy()
    idx = mk == bgval
    out[idx] = bgim[idx]
    return out


def valid_cnstr(x1, y1, x2, y2, msk1, msk2):
    """para_gen.py:216-223"""
    if x1 >= msk1.shape[1] or x2 >= msk2.shape[1] or y1 >= msk1.shape[0] or y2 >= msk2.shape[0]:
        return False
    dist = np.sqrt(float((x2 - x1) ** 2 + (y2 - y1) ** 2))
    return bool(dist < 60 and dist > 0 and msk1[y1, x1] > 0 and msk1[y1, x1] == msk2[y2, x2])


def filter_matches(matches, mk1, mk2):
    """para_gen.py:468-482: keep order; also the label of every kept match (`valids`)."""
    keep, valids = [], []
    for x1, y1, x2, y2 in np.asarray(matches, np.int64).reshape(-1, 4):
        if valid_cnstr(int(x1), int(y1), int(x2), int(y2), mk1, mk2):
            keep.append((x1, y1, x2, y2))
            valids.append(mk1[y1, x1])
    return np.asarray(keep, np.int32).reshape(-1, 4), np.asarray(valids, np.uint8)


def segment_mask(mk1, segment=0, arap_bg=255):
    """para_gen.py:513-527: single-segment mode (segment=0) or one label of --multseg."""
    if segment == 0:
        mask = np.zeros_like(mk1, dtype=np.uint8)
        mask[mk1 == 0] = arap_bg
    else:
        mask = np.zeros_like(mk1, dtype=np.uint8) + arap_bg
        mask[mk1 == segment] = 0
    return mask
