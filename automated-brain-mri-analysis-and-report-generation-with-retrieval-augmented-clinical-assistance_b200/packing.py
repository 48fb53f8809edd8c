"""Weight / activation layout helpers for the tcgen05 conv kernels (host side, torch tensors)."""
import torch


def round_up(x, m):
    return (x + m - 1) // m * m


def pack_conv3_weight(w, cin_pad=None, dtype=torch.bfloat16):
    """nn.Conv3d weight (Cout, Cin, kd, kh, kw) -> 16-bit [27 taps (kd,kw,kh)][cout_pad][cin_pad]."""
    cout, cin = w.shape[0], w.shape[1]
    cin_pad = cin_pad or round_up(cin, 16)
    cout_pad = round_up(cout, 32)
    p = torch.zeros(27, cout_pad, cin_pad, dtype=torch.float32, device=w.device)
    p[:, :cout, :cin] = w.float().permute(2, 4, 3, 0, 1).reshape(27, cout, cin)
    return p.to(dtype).contiguous()


def pack_conv3_weight_kwpacked(w, dtype=torch.bfloat16):
    """First-layer nn.Conv3d weight (Cout, Cin, kd, kh, kw), 3 * Cin <= 16 -> 16-bit [9 taps (kd, kh)][cout_pad][16] with
    K index kw * Cin + ci: the 3x3x1 kernel over the kw-packed input (bsg_gather_patch_tta kwpack = 1)."""
    cout, cin = w.shape[0], w.shape[1]
    assert 3 * cin <= 16
    cout_pad = round_up(cout, 32)
    p = torch.zeros(9, cout_pad, 16, dtype=torch.float32, device=w.device)
    p[:, :cout, :3 * cin] = w.float().permute(2, 3, 0, 4, 1).reshape(9, cout, 3 * cin)  # (kd, kh, co, kw, ci)
    return p.to(dtype).contiguous()


def kwpack_input(x, cpad=16):
    """(N, C, D, H, W) float -> (N, D, H, W, cpad) float with channel k*C + c = x[:, c] shifted by k-1 along w (zero
    outside): the torch-side twin of the gather kernel's kwpack layout (forward_logits convenience path)."""
    n, c, d, h, w = x.shape
    out = torch.zeros(n, d, h, w, cpad, dtype=x.dtype, device=x.device)
    xl = x.permute(0, 2, 3, 4, 1)
    out[..., 1:, 0:c] = xl[..., :-1, :]
    out[..., c:2 * c] = xl
    out[..., :-1, 2 * c:3 * c] = xl[..., 1:, :]
    return out


def pack_conv1_weight(w, cin_pad=None, dtype=torch.bfloat16):
    """1x1x1 conv weight (Cout, Cin, 1, 1, 1) -> 16-bit [1][cout_pad][cin_pad]."""
    cout, cin = w.shape[0], w.shape[1]
    cin_pad = cin_pad or round_up(cin, 16)
    cout_pad = round_up(cout, 32)
    p = torch.zeros(1, cout_pad, cin_pad, dtype=torch.float32, device=w.device)
    p[0, :cout, :cin] = w.float().reshape(cout, cin)
    return p.to(dtype).contiguous()


def pack_convT2_weight(w, cin_pad=None, dtype=torch.bfloat16):
    """nn.ConvTranspose3d k2 s2 weight (Cin, Cout, kd, kh, kw) -> 16-bit [1][8 parities (kd,kh,kw) * cout_pad][cin_pad]."""
    cin, cout = w.shape[0], w.shape[1]
    cin_pad = cin_pad or round_up(cin, 16)
    cout_pad = round_up(cout, 32)
    p = torch.zeros(8, cout_pad, cin_pad, dtype=torch.float32, device=w.device)
    p[:, :cout, :cin] = w.float().permute(2, 3, 4, 1, 0).reshape(8, cout, cin)
    return p.reshape(1, 8 * cout_pad, cin_pad).to(dtype).contiguous()


def split_f16x3(x):
    """fp32 tensor -> (hi, lo) fp16 pair with hi = fp16(x), lo = fp16(x - hi): x = hi + lo to ~22 significant bits."""
    hi = x.float().to(torch.float16)
    return hi, (x.float() - hi.float()).to(torch.float16)


def split_k_weight(w, parts, cin_pad, dim):
    """fp16x3 split of a weight along its input-channel dimension `dim` (engine dtype "fp32"): the source tensor holds
    each part (offset, c) as the three channel blocks [hi | hi | lo], so the weight's K layout is [w_hi | w_lo | w_hi]
    per part (zero elsewhere) and hi*w_hi + hi*w_lo + lo*w_hi = x*w - lo*w_lo.  Values are fp16-exact fp32."""
    hi, lo = split_f16x3(w)
    hi, lo = hi.float(), lo.float()
    shape = list(w.shape)
    shape[dim] = cin_pad
    out = torch.zeros(shape, dtype=torch.float32, device=w.device)
    start = 0
    for off, c in parts:
        h, l = hi.narrow(dim, start, c), lo.narrow(dim, start, c)
        out.narrow(dim, off, c).copy_(h)
        out.narrow(dim, off + c, c).copy_(l)
        out.narrow(dim, off + 2 * c, c).copy_(h)
        start += c
    assert start == w.shape[dim], "parts do not cover the weight's input channels"
    return out


def pad_bias(b, cout):
    cout_pad = round_up(cout, 32)
    out = torch.zeros(cout_pad, dtype=torch.float32, device=b.device if b is not None else None)
    if b is not None:
        out[:cout] = b.float()
    return out


def to_ndhwc_bf16(x, c_pad=None):
    """(N, C, D, H, W) float -> (N, D, H, W, c_pad) bf16, zero-padded channels."""
    n, c, d, h, w = x.shape
    c_pad = c_pad or round_up(c, 16)
    out = torch.zeros(n, d, h, w, c_pad, dtype=torch.bfloat16, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    return out


def from_ndhwc(x, c=None):
    """(N, D, H, W, C) -> (N, c, D, H, W) float32."""
    if c is not None:
        x = x[..., :c]
    return x.permute(0, 4, 1, 2, 3).float().contiguous()
