"""Data ingestion for the registered families from LOCAL files.

The reference gets MNIST through ``tensorflow_datasets`` at import time (models/bayesian_NN/NN_data.py:19-21) - a network
download that has no place on the hot path and no network here.  These loaders read the standard on-disk formats and
return numpy arrays in exactly the layout the model families take (``data`` tuples of gradient_estimation.py:25-29):

* ``load_mnist_idx``      IDX files (``train-images-idx3-ubyte[.gz]`` + labels) -> (X float32 [N, 784] in [0, 1],
                          y float32 [N, 10] one-hot), the arrays NN_data.py:26-42 builds (pixels / 255, one-hot float labels).
* ``load_movielens_100k`` ``u.data`` (tab separated ``user item rating timestamp``, 1-BASED ids) -> (ij int32 [N, 2] 0-based,
                          r float32 [N]) + (n_users, n_items): the PMF family's data tuple.  The native
                          library refuses out-of-range ids at ingestion, so forgetting the -1 is an error, not corruption.
* ``one_hot``             NN_data.py:26-31.
"""
from __future__ import annotations

import gzip
import struct
from typing import Tuple

import numpy as np


def one_hot(labels, n_classes: int) -> np.ndarray:
    """float32 one-hot rows (NN_data.py:26-31)."""
    labels = np.asarray(labels).astype(np.int64).ravel()
    if labels.size and (labels.min() < 0 or labels.max() >= n_classes):
        raise ValueError("label outside [0, n_classes)")
    out = np.zeros((labels.size, n_classes), dtype=np.float32)
    out[np.arange(labels.size), labels] = 1.0
    return out


def _open(path: str):
    return gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")


def read_idx(path: str) -> np.ndarray:
    """One IDX file (the MNIST container): magic 0x0000 <dtype> <ndim>, big-endian dims, raw data."""
    with _open(path) as f:
        zero, dtype_code, ndim = struct.unpack(">HBB", f.read(4))
        if zero != 0:
            raise ValueError(f"{path}: not an IDX file")
        dtypes = {0x08: np.uint8, 0x09: np.int8, 0x0B: ">i2", 0x0C: ">i4", 0x0D: ">f4", 0x0E: ">f8"}
        if dtype_code not in dtypes:
            raise ValueError(f"{path}: unknown IDX dtype code {dtype_code:#x}")
        dims = struct.unpack(">" + "I" * ndim, f.read(4 * ndim))
        data = np.frombuffer(f.read(), dtype=dtypes[dtype_code])
    if data.size != int(np.prod(dims)):
        raise ValueError(f"{path}: {data.size} elements, header says {dims}")
    return data.reshape(dims)


def load_mnist_idx(images_path: str, labels_path: str, n_classes: int = 10) -> Tuple[np.ndarray, np.ndarray]:
    """-> (X float32 [N, rows * cols] scaled to [0, 1], y float32 [N, n_classes] one-hot): NN_data.py:33-45."""
    images = read_idx(images_path)
    labels = read_idx(labels_path)
    if images.shape[0] != labels.shape[0]:
        raise ValueError("image and label files disagree on N")
    X = (images.reshape(images.shape[0], -1).astype(np.float32) / np.float32(255.0)).astype(np.float32)
    return X, one_hot(labels, n_classes)


def load_movielens_100k(path: str):
    """``u.data`` -> ((ij int32 [N, 2] ZERO-based, r float32 [N]), (n_users, n_items))."""
    raw = np.loadtxt(path, dtype=np.int64, usecols=(0, 1, 2), ndmin=2)
    if raw.size == 0:
        raise ValueError(f"{path}: no ratings")
    if raw[:, :2].min() < 1:
        raise ValueError(f"{path}: MovieLens ids are 1-based; found an id < 1")
    ij = (raw[:, :2] - 1).astype(np.int32)
    r = raw[:, 2].astype(np.float32)
    return (np.ascontiguousarray(ij), r), (int(ij[:, 0].max()) + 1, int(ij[:, 1].max()) + 1)
