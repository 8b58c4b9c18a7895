"""Drop-in for CRF/krahenbuhl2013/CRF.py: same function, same arguments, same return value."""
from krahenbuhl2013.wrapper import DenseCRF

__all__ = ["CRF", "DenseCRF"]


def CRF(image, unary, maxiter=10, scale_factor=1.0, color_factor=13):
    """Mean-field inference in a fully connected CRF with Gaussian edge potentials.

    image : (H, W, 3) array, values in [0, 256) (cast to ubyte like the reference, CRF.py:32)
    unary : (H, W, M) array; the energies are ``-unary`` (CRF.py:28)
    Returns the (H, W, M) float32 marginals after ``maxiter`` iterations (CRF.py:35-37).
    """
    assert(image.shape[:2] == unary.shape[:2])
    H, W = image.shape[:2]
    nlables = unary.shape[2]
    crf = DenseCRF(W, H, nlables)
    crf.set_unary_energy(-unary.ravel().astype('float32'))
    crf.add_pairwise_energy(10, 80 / scale_factor, 80 / scale_factor, color_factor, color_factor, color_factor,
                            3, 3 / scale_factor, 3 / scale_factor, image.ravel().astype('ubyte'))
    prediction = crf.inference(maxiter).reshape((H, W, nlables))
    return prediction
