"""numpy restatement of the reference's layer flatten + background composite (TEST INFRASTRUCTURE ONLY).

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
