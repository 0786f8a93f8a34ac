"""Random-init weights in the reference checkpoints' key layout, for synthetic benchmarks and
tests when no checkpoint is available.  Keras ``HeNormal`` kernels (m4depth_network.py:61,100: stddev sqrt(2/fan_in);
a plain normal is used instead of Keras' truncated normal), zero biases, DN scale 1 / bias 0 (:35-38)."""
import math

import torch

ENC_CHANNELS = [16, 32, 64, 96, 128, 192]          # m4depth_network.py:59
PREP_CHANNELS = [128, 128, 96]                     # :102
EST_CHANNELS = [64, 32, 16, 5]                     # :109


def refiner_in_channels(lvl_depth):
    cuts = 2 ** (lvl_depth // 2)
    return 9 * cuts + 1 + 4 + 49 * cuts + 1          # :223-242


def init_random_weights(nbre_levels=6, seed=7, bias_std=0.0):
    g = torch.Generator().manual_seed(seed)
    w = {}

    def conv(name, cin, cout):
        w[name + "/kernel"] = torch.randn(3, 3, cin, cout, generator=g, dtype=torch.float32) * math.sqrt(2.0 / (9 * cin))
        w[name + "/bias"] = torch.randn(cout, generator=g, dtype=torch.float32) * bias_std

    cin = 3
    for i in range(nbre_levels):
        conv(f"encoder/conv_layers_s1/{i}", cin, ENC_CHANNELS[i])
        conv(f"encoder/conv_layers_s2/{i}", ENC_CHANNELS[i], ENC_CHANNELS[i])
        cin = ENC_CHANNELS[i]
    w["encoder/dn_layers/0/scale"] = torch.ones(1, 1, 1, ENC_CHANNELS[0], dtype=torch.float32)
    w["encoder/dn_layers/0/bias"] = torch.zeros(1, 1, 1, ENC_CHANNELS[0], dtype=torch.float32)
    for lvl in range(1, nbre_levels + 1):
        cin = refiner_in_channels(lvl)
        p = f"d_estimator/levels/{lvl - 1}/disp_refiner"
        for j, co in enumerate(PREP_CHANNELS):
            conv(f"{p}/prep_conv_layers/{j}", cin, co)
            cin = co
        for j, co in enumerate(EST_CHANNELS):
            conv(f"{p}/est_d_conv_layers/{j}", cin, co)
            cin = co
    return w
