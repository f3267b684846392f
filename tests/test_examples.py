"""CPU test of the example driver's data recipe (examples/synthetic_data_hard.py): shape and standardisation of
BASELINE.json configs[1] (test/synthetic_data_hard_test.py: 100 x 60, four groups of 15 GP draws)."""
import importlib.util
import os

import numpy as np

from conftest import ROOT


def test_hard_synthetic_data_recipe():
    spec = importlib.util.spec_from_file_location("synthetic_data_hard", os.path.join(ROOT, "examples", "synthetic_data_hard.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    y, labels = mod.make_data()
    assert y.shape == (100, 60) and labels.shape == (60,)
    assert np.array_equal(np.bincount(labels), [15, 15, 15, 15])
    assert np.abs(y.mean(0)).max() < 1e-12 and np.abs(y.std(0) - 1.0).max() < 1e-12
    y2, _ = mod.make_data()
    assert np.array_equal(y, y2)                       # seeded
    # columns of one group share their inputs: they correlate more within the group than across groups on average
    c = np.abs(np.corrcoef(y.T))
    same = labels[:, None] == labels[None, :]
    off = ~np.eye(60, dtype=bool)
    assert c[same & off].mean() > c[~same].mean()
