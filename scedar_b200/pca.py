"""PCA / truncated-SVD producer of the kNN path.

Placeholder module body replaced below once the device kernels exist.
"""


def pca_x(x):
    raise NotImplementedError(
        "PCA-space kNN needs scedar's PCA (sdm.py:1313-1334); pass the PCs "
        "as x instead")
