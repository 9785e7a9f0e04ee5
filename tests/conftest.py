import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

MODEL_CASES = {
    # fixture name -> (kind, scorer, loss, smoothing, pool, batchnorm, optimizer)
    "lookup_distmult_bce": ("lookup", "distmult", "bce", 0.0, "sum", False, "adagrad"),
    "lookup_complex_bce": ("lookup", "complex", "bce", 0.0, "sum", False, "adagrad"),
    "lookup_complex_bce_smooth": ("lookup", "complex", "bce", 0.1, "sum", False, "adagrad"),
    "lookup_complex_kl_adam": ("lookup", "complex", "kl", 0.0, "sum", False, "adam"),
    "unigram_complex_sum_bce": ("unigram", "complex", "bce", 0.0, "sum", False, "adagrad"),
    "unigram_complex_mean_bce": ("unigram", "complex", "bce", 0.0, "mean", False, "adagrad"),
    "unigram_complex_max_bce": ("unigram", "complex", "bce", 0.0, "max", False, "adagrad"),
    "unigram_complex_sum_bn_bce": ("unigram", "complex", "bce", 0.0, "sum", True, "adagrad"),
    "lstm_complex_bce": ("lstm", "complex", "bce", 0.0, "sum", False, "adagrad"),
    "lstm_distmult_bn_bce": ("lstm", "distmult", "bce", 0.0, "sum", True, "adagrad"),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def kats():
    return load_golden("kats")


@pytest.fixture(scope="session", params=sorted(MODEL_CASES))
def model_case(request):
    return request.param, MODEL_CASES[request.param], load_golden(request.param)


def params_of(gold, prefix):
    return {k[len(prefix):]: v for k, v in gold.items() if k.startswith(prefix)}
