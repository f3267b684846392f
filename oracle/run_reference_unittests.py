"""TEST INFRASTRUCTURE ONLY -- runs the reference's OWN unit tests (test/unittests/*.py, unmodified)
against the reference's own model code executed over the oracle's TensorFlow-1 shim.

Purpose: validate the shim.  Each reference test compares the reference graph against the reference's
pure-numpy "naive" re-computation (rtol 1e-7); if they pass here, the shim reproduces TF-1.15 semantics
for every op on the DP-GP-LVM path, and fixtures generated through it (oracle/make_golden.py) can be
trusted as outputs of the reference itself.

Skipped: TestFasterDPGPLVM (trains two models for 5000 Adam steps through `tf.train`, which the shim does
not provide; its initial-objective equality check is reproduced in oracle/make_golden.py).

Usage (this container only; needs /root/reference):  python -m oracle.run_reference_unittests
"""
import sys
import unittest


def main():
    from oracle import ref_env
    ref_env.activate()
    names = ["test.unittests.kernel_unittests", "test.unittests.dp_unittests",
             "test.unittests.bgplvm_unittests", "test.unittests.dpgplvm_unitttests"]
    loader = unittest.TestLoader()
    suite = unittest.TestSuite()
    for n in names:
        mod = __import__(n, fromlist=["x"])
        for attr in dir(mod):
            obj = getattr(mod, attr)
            if isinstance(obj, type) and issubclass(obj, unittest.TestCase) and obj.__module__ == n:
                if attr == "TestFasterDPGPLVM":
                    continue
                suite.addTests(loader.loadTestsFromTestCase(obj))
    res = unittest.TextTestRunner(verbosity=2).run(suite)
    return 0 if res.wasSuccessful() else 1


if __name__ == "__main__":
    sys.exit(main())
