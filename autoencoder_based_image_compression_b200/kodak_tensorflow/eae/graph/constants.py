"""Architecture constants of the entropy autoencoder on the inference path.

Values follow kodak_tensorflow/eae/graph/constants.py:26-59 of the reference (only the ones the
analysis / synthesis transforms and the quantizer use; the training learning rates are out of scope).
"""

# Lower bound of the GDN/IGDN weights and additive coefficients (constants.py:26).
MIN_GAMMA_BETA = 2.e-5

# Projection bounds of the learned quantization bin widths (constants.py:30-31).
MIN_BW = 0.8
MAX_BW = 4.

NB_MAPS_1 = 128
NB_MAPS_2 = 128
NB_MAPS_3 = 128
WIDTH_KERNEL_1 = 9
WIDTH_KERNEL_2 = 5
WIDTH_KERNEL_3 = 5
STRIDE_1 = 4
STRIDE_2 = 2
STRIDE_3 = 2

# Ratio between the image size and the latent map size (constants.py:59).
STRIDE_PROD = STRIDE_1*STRIDE_2*STRIDE_3
