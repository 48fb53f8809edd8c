"""Seeded synthetic inputs and random-init BraTS-2021-shaped models shared by bench.py, the tests,
smoke() and the oracle's golden-vector scripts.  Plain numpy / torch / scipy data generation: no reference code, no
oracle code, nothing here computes a result that is being checked."""
import os

import numpy as np
import torch
from scipy.ndimage import gaussian_filter
from torch import nn

ROOT = os.path.dirname(os.path.abspath(__file__))


def case_volume(seed=0, shape=(4, 155, 240, 240)):
    """BASELINE config 1/2 input: randn(4,155,240,240) fp32, array order (C, z, y, x)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).numpy()


def label_volume(seed=0, shape=(240, 240, 155), sigma=6.0):
    """Blobby label volume: smoothed noise thresholded at 1.5/2.0/2.5 sigma -> labels 1/2/3 (≈6 % tumour)."""
    rng = np.random.default_rng(seed)
    f = gaussian_filter(rng.standard_normal(shape), sigma)
    f = (f - f.mean()) / f.std()
    lab = np.zeros(shape, dtype=np.uint8)
    lab[f > 1.5] = 1
    lab[f > 2.0] = 2
    lab[f > 2.5] = 3
    return lab


def label_pair(seed=0, shape=(240, 240, 155)):
    """(prediction, ground truth) pair: the prediction is the GT rolled by 3 voxels along axis 0."""
    gt = label_volume(seed, shape)
    return np.roll(gt, 3, axis=0).copy(), gt


def mri_volumes(seed, seg):
    """Four MRI-like modalities for a label volume: an ellipsoidal "head" of smooth positive texture (zero outside),
    the tumour labels scale the signal per modality.  Integer-valued float32 (like int16 NIfTI data), so order
    statistics meet ties and every value is exact in float32 and float64."""
    rng = np.random.default_rng(1000 + seed)
    shape = seg.shape
    grids = np.meshgrid(*[np.linspace(-1.0, 1.0, s) for s in shape], indexing="ij")
    head = sum(g ** 2 for g in grids) < 2.2
    gains = {"t1": (0.6, 1.0, 0.9), "t1ce": (0.7, 1.0, 1.8), "t2": (1.9, 1.5, 1.2), "flair": (0.8, 1.6, 1.3)}
    out = {}
    for name, per_label in gains.items():
        tex = gaussian_filter(rng.standard_normal(shape), 2.0)
        tex = 400.0 + 120.0 * tex / tex.std()
        factor = np.ones(shape)
        for lab, gain in zip((1, 2, 3), per_label):
            factor[seg == lab] = gain
        out[name] = (np.round(np.clip(tex * factor, 1.0, None)) * head).astype(np.float32)
    return out


def randomize_norm_params(net, seed):
    """Random-init leaves every norm at weight=1, bias=0, running stats (0,1): randomise them (seeded) so that BN
    folding and the affine terms are actually exercised by the parity tests."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, (nn.BatchNorm3d, nn.GroupNorm, nn.InstanceNorm3d)):
                if m.weight is not None:
                    m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=g))
                    m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                if isinstance(m, nn.BatchNorm3d):
                    m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                    m.running_var.copy_(1.0 + 0.3 * torch.rand(m.running_var.shape, generator=g))
            if isinstance(m, nn.Conv3d) and m.bias is not None:
                m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=g))


def build_dropin_unet(variant="bn", base=32, num_pool=5, in_ch=4, num_classes=3, seed=1, groups=8, encoder_scale=1,
                      max_num_features=None, nonlin="sigmoid"):
    """The drop-in Generic_UNet built the way the BraTS-2021 V2 trainers build the reference class."""
    from brainseg_b200 import generic_UNet as G

    norm_op = {"bn": nn.BatchNorm3d, "in": nn.InstanceNorm3d, "gn": nn.GroupNorm}[variant]
    norm_kwargs = {"eps": 1e-5, "affine": True}
    if variant == "gn":
        norm_kwargs["num_groups"] = groups
    torch.manual_seed(seed)
    net = G.Generic_UNet(in_ch, base, num_classes, num_pool, 2, 2, nn.Conv3d, norm_op, norm_kwargs, nn.Dropout3d,
                         {"p": 0, "inplace": True}, nn.LeakyReLU, {"negative_slope": 1e-2, "inplace": True}, True,
                         False, lambda x: x, G.InitWeights_He(1e-2), [[2, 2, 2]] * num_pool,
                         [[3, 3, 3]] * (num_pool + 1), False, True, True, max_num_features=max_num_features,
                         encoder_scale=encoder_scale)
    randomize_norm_params(net, seed + 1000)
    net.eval()
    net.do_ds = False
    if nonlin == "sigmoid":
        net.inference_apply_nonlin = nn.Sigmoid()  # BraTSRegions trainers
    elif nonlin == "softmax":
        net.inference_apply_nonlin = G.softmax_helper
    return net


def build_benchmark_models(model2="large"):
    """The two ensemble members of BASELINE configs[1]: model 1 = Generic_UNet BatchNorm (31.2 M parameters), model 2 =
    the large GroupNorm variant (87.4 M; encoder_scale 2, max 512 features) or the standard-size GroupNorm net."""
    m1 = build_dropin_unet("bn", base=32, num_pool=5, seed=1)
    if model2 == "large":
        m2 = build_dropin_unet("gn", base=32, num_pool=5, seed=2, groups=8, encoder_scale=2, max_num_features=512)
    elif model2 == "standard":
        m2 = build_dropin_unet("gn", base=32, num_pool=5, seed=2, groups=8)
    else:
        raise ValueError(model2)
    return m1, m2


# ---------------------------------------------------------------------------------------------------------------
# tests/golden/config2_oracle.npz: BASELINE configs[1] in full through the CPU oracle (oracle/make_config2_golden.py)
# ---------------------------------------------------------------------------------------------------------------
CONFIG2_ORACLE = os.path.join(ROOT, "tests", "golden", "config2_oracle.npz")


def pack2(lab):
    """uint8 labels 0..3 -> 2 bits per voxel."""
    flat = np.asarray(lab).reshape(-1).astype(np.uint8)
    pad = (-flat.size) % 4
    flat = np.concatenate([flat, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (flat[:, 0] | (flat[:, 1] << 2) | (flat[:, 2] << 4) | (flat[:, 3] << 6)).astype(np.uint8)


def unpack2(packed, shape):
    n = int(np.prod(shape))
    out = np.empty((packed.size, 4), np.uint8)
    for k in range(4):
        out[:, k] = (packed >> (2 * k)) & 3
    return out.reshape(-1)[:n].reshape(shape)


def load_config2_oracle(path=CONFIG2_ORACLE):
    """-> dict: seg1, seg2 (uint8 (z, y, x) oracle label volumes of the two models), decisive1/2 (bool: every class
    probability further than `tol` from 0.5), probs1/2 (float32 class probabilities on the lattice [:, ::L, ::L, ::L]),
    lattice, tol.  None when the file is absent."""
    if not os.path.exists(path):
        return None
    z = np.load(path)
    shape = tuple(int(s) for s in z["shape"])
    n = int(np.prod(shape))
    out = {"shape": shape, "lattice": int(z["lattice"]), "tol": float(z["tol"])}
    for m in (1, 2):
        out[f"seg{m}"] = unpack2(z[f"seg{m}"], shape)
        out[f"decisive{m}"] = np.unpackbits(z[f"decisive{m}"])[:n].astype(bool).reshape(shape)
        out[f"probs{m}"] = z[f"probs{m}"]
    return out
