"""Shared fixtures-from-goldens helpers (tests only)."""
import os

import numpy as np

from tracktolearn_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def subject_for(g, vol_seed=1234):
    shape = tuple(int(v) for v in g['meta_shape'])
    sub = synthetic.make_subject(shape, seed=vol_seed)
    return {k: (v.numpy() if v is not None else None) for k, v in sub.items()}


def meta(g):
    vox, step_mm, theta, max_length, thr, max_nb_steps, step_vox = [float(v) for v in g['meta']]
    return dict(vox=vox, step_mm=step_mm, theta=theta, max_length=max_length, threshold=thr,
                max_nb_steps=int(max_nb_steps), step_vox=step_vox)


def split_by_counts(arr, counts):
    out, o = [], 0
    for c in counts:
        out.append(arr[o:o + c])
        o += c
    return out


def oracle_ckpt_for(g):
    """The 1-layer TransformerOracle checkpoint of the env_oracle fixture (head rescaled so that
    scores straddle 0.5; the rescaled head travels in the fixture)."""
    import torch
    from tracktolearn_b200 import synthetic
    ck = synthetic.oracle_checkpoint(n_head=4, n_layers=1, input_size=384, seed=31)
    ck['state_dict']['head.weight'] = torch.from_numpy(np.asarray(g['head_w']))
    ck['state_dict']['head.bias'] = torch.from_numpy(np.asarray(g['head_b']))
    return ck
