"""Shared helpers for the test-suite: golden fixtures + key-seeded synthetic weights."""
import json
import os

import numpy as np
import torch

import dsdiff_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def manifest():
    with open(os.path.join(GOLD, "manifest.json")) as f:
        return json.load(f)


def synth_sd(case):
    m = manifest()[case]
    return m["hp"], O.synth_state_dict(m["shapes"])


def maxabs(a, b):
    return float((a.double() - b.double()).abs().max())
