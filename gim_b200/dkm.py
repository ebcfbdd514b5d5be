"""Drop-in for the reference's DKMv3 dense matcher on B200.

`DKMv3(weights, h, w, symmetric=True, sample_mode="threshold_balanced", **kw)` keeps the contract of
`networks/dkm/models/model_zoo/DKMv3.py:5` / `networks/dkm/models/dkm.py:537-752`: it returns an `nn.Module` whose
attributes `h_resized, w_resized, upsample_preds, upsample_res, symmetric, sample_thresh,
use_soft_mutual_nearest_neighbours` callers overwrite after construction (`trainer/lightning.py:32-37`), with the same
`state_dict` keys (`encoder.net.*`, `decoder.*`), `.match(im1, im2) -> (warp [H, 2W, 4], certainty [H, 2W])` and
`.sample(warp, certainty, num) -> (matches [n, 4], confidence [n])`.

match() runs entirely in libgimb200.so (csrc/dkm_api.cu: tcgen05 GEMM engine + the kernels of dkm_kernels.cu); there is
no PyTorch forward and no CPU path.  sample() stays in torch with the caller's RNG, exactly as SURVEY.md section 8 (a2.8)
prescribes: its output depends on `torch.multinomial`, so parity for this model is defined on match()'s dense output."""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .dkm_params import DKMParams, REFINER_CFG
from .weights import _ENTRY, _HEADER, BLOB_MAGIC, _fold_bn, _khwc


# ----------------------------------------------------------------------------- packed blob for gimb_dkm_create
def packed_dkm_tensors(sd):
    """state_dict (reference key names) -> ordered {packed name: fp32 tensor} consumed by csrc/dkm_api.cu::build_dkm."""
    out = {}
    sd = {k: v.detach().cpu() for k, v in sd.items()}  # packing happens on the host
    f = lambda k: sd[k].float()

    def conv_bn(name, wkey, bnkey):  # bias-free conv + BatchNorm (torchvision ResNet)
        out[name + ".w"] = _khwc(sd[wkey])
        out[name + ".s"], out[name + ".b"] = _fold_bn(sd, bnkey)

    def conv_bias(name, pre, pad_to=None, bn=None):  # conv with bias (+ optional BatchNorm after it)
        w, b = _khwc(sd[pre + ".weight"]), f(pre + ".bias")
        s = torch.ones_like(b)
        if bn is not None:  # y = bn_s * (conv + bias) + bn_b
            bs, bb = _fold_bn(sd, bn)
            s, b = bs, bs * b + bb
        if pad_to is not None and w.shape[0] < pad_to:  # heads with 3 outputs: the GEMM engine wants >= 8 columns
            n = pad_to - w.shape[0]
            w = torch.cat([w, torch.zeros(n, *w.shape[1:])])
            s = torch.cat([s, torch.ones(n)])
            b = torch.cat([b, torch.zeros(n)])
        out[name + ".w"], out[name + ".s"], out[name + ".b"] = w.contiguous(), s.contiguous(), b.contiguous()

    e = "encoder.net"
    conv_bn("enc.stem", e + ".conv1.weight", e + ".bn1")
    for li, nblk in ((1, 3), (2, 4), (3, 6), (4, 3)):
        for bi in range(nblk):
            pre, name = f"{e}.layer{li}.{bi}", f"enc.l{li}.{bi}"
            for ci in (1, 2, 3):
                conv_bn(f"{name}.c{ci}", f"{pre}.conv{ci}.weight", f"{pre}.bn{ci}")
            if f"{pre}.downsample.0.weight" in sd:
                conv_bn(f"{name}.ds", f"{pre}.downsample.0.weight", f"{pre}.downsample.1")
    d = "decoder"
    for s in ("32", "16"):
        conv_bias(f"proj.{s}", f"{d}.proj.{s}")
        out[f"gp.{s}.pos_w"] = f(f"{d}.gps.{s}.pos_conv.weight").reshape(-1, 2).contiguous()
        out[f"gp.{s}.pos_b"] = f(f"{d}.gps.{s}.pos_conv.bias")
        ed = f"{d}.embedding_decoder"
        conv_bias(f"dfn.{s}.feat", f"{ed}.feat_input_modules.{s}")
        for short, mod in (("rrbd", "rrb_d"), ("rrbu", "rrb_u")):
            conv_bias(f"dfn.{s}.{short}.c1", f"{ed}.{mod}.{s}.conv1")
            conv_bias(f"dfn.{s}.{short}.c2", f"{ed}.{mod}.{s}.conv2", bn=f"{ed}.{mod}.{s}.bn")
            conv_bias(f"dfn.{s}.{short}.c3", f"{ed}.{mod}.{s}.conv3")
        out[f"dfn.{s}.cab.w1"] = f(f"{ed}.cab.{s}.conv1.weight").reshape(384, 768).contiguous()
        out[f"dfn.{s}.cab.b1"] = f(f"{ed}.cab.{s}.conv1.bias")
        out[f"dfn.{s}.cab.w2"] = f(f"{ed}.cab.{s}.conv2.weight").reshape(384, 384).contiguous()
        out[f"dfn.{s}.cab.b2"] = f(f"{ed}.cab.{s}.conv2.bias")
        conv_bias(f"dfn.{s}.term", f"{ed}.terminal_module.{s}", pad_to=8)
    for s in REFINER_CFG:
        r = f"{d}.conv_refiner.{s}"
        out[f"ref.{s}.emb_w"] = f(r + ".disp_emb.weight").reshape(-1, 2).contiguous()
        out[f"ref.{s}.emb_b"] = f(r + ".disp_emb.bias")
        for k in range(9):
            blk = f"{r}.block1" if k == 0 else f"{r}.hidden_blocks.{k - 1}"
            w = f(blk + ".0.weight")  # [C, 1, 5, 5] depthwise
            bs, bb = _fold_bn(sd, blk + ".1")
            out[f"ref.{s}.b{k}.dw_w"] = w.reshape(w.shape[0], 25).contiguous()
            out[f"ref.{s}.b{k}.dw_s"] = bs.contiguous()
            out[f"ref.{s}.b{k}.dw_b"] = (bs * f(blk + ".0.bias") + bb).contiguous()
            cp = (w.shape[0] + 31) // 32 * 32  # transposed + zero-padded copies for the vectorised depthwise kernel
            wt = torch.zeros(25, cp); wt[:, : w.shape[0]] = w.reshape(w.shape[0], 25).t()
            sp = torch.zeros(cp); sp[: w.shape[0]] = out[f"ref.{s}.b{k}.dw_s"]
            bp = torch.zeros(cp); bp[: w.shape[0]] = out[f"ref.{s}.b{k}.dw_b"]
            out[f"ref.{s}.b{k}.dw_wt"], out[f"ref.{s}.b{k}.dw_sp"], out[f"ref.{s}.b{k}.dw_bp"] = wt, sp, bp
            conv_bias(f"ref.{s}.b{k}.pw", blk + ".3")
        conv_bias(f"ref.{s}.out", r + ".out_conv", pad_to=8)
    return out


def pack_dkm_blob(state_dict):
    """-> bytes: the blob `gimb_dkm_create` takes (container layout of include/gimb200.h)."""
    tensors = packed_dkm_tensors(state_dict)
    entries, chunks, off = [], [], 0
    for name, t in tensors.items():
        a = t.detach().cpu().contiguous().numpy().astype(np.float32, copy=False)
        shape = list(a.shape) + [0] * (4 - a.ndim)
        entries.append(_ENTRY.pack(name.encode(), a.ndim, *shape, 0, off, a.nbytes))
        raw = a.tobytes()
        pad = (-len(raw)) % 256
        chunks.append(raw + b"\0" * pad)
        off += len(raw) + pad
    table = b"".join(entries)
    data_offset = (_HEADER.size + len(table) + 255) // 256 * 256
    head = _HEADER.pack(BLOB_MAGIC, 1, len(entries), data_offset, data_offset + off)
    return head + table + b"\0" * (data_offset - _HEADER.size - len(table)) + b"".join(chunks)


class _Taps(ctypes.Structure):
    _fields_ = [("enc", ctypes.c_void_p * 6), ("gp32", ctypes.c_void_p), ("gp16", ctypes.c_void_p), ("flow", ctypes.c_void_p * 6),
                ("cert", ctypes.c_void_p * 6), ("flow_up", ctypes.c_void_p * 6), ("cert_up", ctypes.c_void_p * 6),
                ("dfn_flow16", ctypes.c_void_p), ("refiner_in16", ctypes.c_void_p), ("refiner_dw16", ctypes.c_void_p),
                ("refiner_pw16", ctypes.c_void_p), ("refiner_out16", ctypes.c_void_p)]


class RegressionMatcher(DKMParams):
    """networks/dkm/models/dkm.py:537-752 on libgimb200 (parameters: `encoder.net.*`, `decoder.*`)."""

    def __init__(self, h=384, w=512, sample_mode="threshold", upsample_preds=True, symmetric=False, name=None,
                 use_soft_mutual_nearest_neighbours=False, **_ignored):
        super().__init__()
        self.w_resized, self.h_resized = w, h
        self.sample_mode = sample_mode
        self.upsample_preds = upsample_preds
        self.symmetric = symmetric
        self.name = name
        self.sample_thresh = 0.05
        self.upsample_res = (1152, 1536)
        if use_soft_mutual_nearest_neighbours and not symmetric:
            raise AssertionError("MNS requires symmetric inference")
        self.use_soft_mutual_nearest_neighbours = use_soft_mutual_nearest_neighbours
        self._handle = None
        self._handle_device = None
        self._ws = None
        self.debug_taps = None  # set to a list of tap names to receive them in self.last_taps (tests)

    # ---- handle management -------------------------------------------------------------------
    def _drop(self):
        if self._handle is not None:
            _lib.load().gimb_dkm_destroy(self._handle)
        self._handle, self._handle_device, self._ws = None, None, None

    def __del__(self):
        try:
            self._drop()
        except Exception:
            pass

    def load_state_dict(self, state_dict, strict=True):
        r = super().load_state_dict(state_dict, strict=strict)
        self._drop()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._drop()
        return r

    def _ensure(self, device):
        if device.type != "cuda":
            raise RuntimeError("gim_b200.DKMv3 has no CPU path: inputs must be CUDA tensors on a B200 (sm_100a)")
        dev = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._handle_device == dev:
            return dev
        self._drop()
        lib = _lib.load()
        blob = pack_dkm_blob(self.state_dict())
        h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.gimb_dkm_create(blob, len(blob), dev, ctypes.byref(h)))
        self._handle, self._handle_device = h, dev
        return dev

    def launch_count(self):
        return int(_lib.load().gimb_dkm_launch_count(self._handle)) if self._handle else 0

    # ---- RegressionMatcher.match (dkm.py:655-752) ---------------------------------------------
    @torch.no_grad()
    def match(self, im1_path, im2_path, *args, batched=False):
        if batched:
            raise NotImplementedError("batched=True is not used by any gim caller (demo.py, trainer/lightning.py, hloc)")
        if not self.symmetric:
            raise NotImplementedError("gim constructs DKMv3 with symmetric=True (model_zoo/DKMv3.py:5); the asymmetric branch of the "
                                      "reference raises in its upsample pass (dkm.py:702)")
        im1, im2 = im1_path, im2_path
        if im1.dim() != 4 or im1.shape[0] != 1 or im1.shape[1] != 3 or im2.dim() != 4 or im2.shape[0] != 1 or im2.shape[1] != 3:
            raise RuntimeError(f"match expects two [1, 3, H, W] tensors, got {tuple(im1.shape)} and {tuple(im2.shape)}")
        dev = self._ensure(im1.device)
        lib = _lib.load()
        im1 = im1.to(torch.float32).contiguous()
        im2 = im2.to(device=im1.device, dtype=torch.float32).contiguous()
        H1, W1, H2, W2 = im1.shape[2], im1.shape[3], im2.shape[2], im2.shape[3]
        h, w = int(self.h_resized), int(self.w_resized)
        up = 1 if self.upsample_preds else 0
        uh, uw = (int(self.upsample_res[0]), int(self.upsample_res[1])) if up else (h, w)
        need = ctypes.c_size_t()
        with torch.cuda.device(dev):
            _lib.check(lib.gimb_dkm_workspace_bytes(self._handle, H1, W1, H2, W2, h, w, up, uh, uw, ctypes.byref(need)))
            if self._ws is None or self._ws.numel() < need.value:
                self._ws = None
                self._ws = torch.empty(need.value, dtype=torch.uint8, device=im1.device)
            ho, wo = (uh, uw) if up else (h, w)
            warp = torch.empty(ho, 2 * wo, 4, dtype=torch.float32, device=im1.device)
            cert = torch.empty(ho, 2 * wo, dtype=torch.float32, device=im1.device)
            taps_ptr, self.last_taps = None, {}
            if self.debug_taps:
                taps = _Taps()
                for name in self.debug_taps:
                    if name in ("dfn_flow16", "refiner_in16", "refiner_dw16", "refiner_pw16", "refiner_out16"):
                        h16, w16 = h, w
                        for _ in range(4):
                            h16, w16 = (h16 + 1) // 2, (w16 + 1) // 2
                        t = torch.zeros(2, h16, w16, 2 if name == "dfn_flow16" else 1377, device=im1.device)
                        setattr(taps, name, t.data_ptr())
                        self.last_taps[name] = t
                        continue
                    kind, s = name.rstrip("u").rstrip("0123456789"), None
                    digits = "".join(ch for ch in name if ch.isdigit())
                    s = int(digits).bit_length() - 1 if digits else 0
                    upass = name.endswith("u")
                    hs, ws = ((uh, uw) if upass else (h, w))
                    for _ in range(s):                       # every stride-2 stage maps n -> ceil(n / 2)
                        hs, ws = (hs + 1) // 2, (ws + 1) // 2
                    if kind == "enc":
                        t = torch.zeros(2, hs, ws, (3, 64, 256, 512, 1024, 2048)[s], device=im1.device)
                        taps.enc[s] = t.data_ptr()
                    elif kind == "gp":
                        t = torch.zeros(2, hs, ws, 256, device=im1.device)
                        setattr(taps, name, t.data_ptr())
                    elif kind == "flow":
                        t = torch.zeros(2, hs, ws, 2, device=im1.device)
                        (taps.flow_up if upass else taps.flow)[s] = t.data_ptr()
                    elif kind == "cert":
                        t = torch.zeros(2, hs, ws, device=im1.device)
                        (taps.cert_up if upass else taps.cert)[s] = t.data_ptr()
                    else:
                        raise KeyError(name)
                    self.last_taps[name] = t
                taps_ptr = ctypes.byref(taps)
            _lib.check(lib.gimb_dkm_match(self._handle, im1.data_ptr(), H1, W1, im2.data_ptr(), H2, W2, h, w, up, uh, uw,
                                          self._ws.data_ptr(), self._ws.numel(), warp.data_ptr(), cert.data_ptr(), taps_ptr,
                                          torch.cuda.current_stream(im1.device).cuda_stream))
        return warp, cert

    # ---- RegressionMatcher.sample (dkm.py:583-620) + kde (utils/kde.py:17-26): torch, caller's RNG ----
    def sample(self, dense_matches, dense_certainty, num=10000):
        certainty_raw = dense_certainty.reshape(-1)
        if "threshold" in self.sample_mode:
            certainty = dense_certainty.clone()
            certainty[certainty > self.sample_thresh] = 1
        elif "pow" in self.sample_mode:
            certainty = dense_certainty ** (1 / 3)
        elif "naive" in self.sample_mode:
            certainty = torch.ones_like(dense_certainty)
        else:
            certainty = dense_certainty
        matches, certainty = dense_matches.reshape(-1, 4), certainty.reshape(-1)
        balanced = "balanced" in self.sample_mode
        if not certainty.sum():
            certainty = certainty + 1e-8
        good = torch.multinomial(certainty, num_samples=min((4 if balanced else 1) * num, len(certainty)), replacement=False)
        good_matches = matches[good]
        # the reference reports the un-thresholded certainty of the drawn matches (`good_certainty = good_certainty_`)
        good_certainty = certainty_raw[good] if "threshold" in self.sample_mode else certainty[good]
        if not balanced:
            return good_matches, good_certainty
        density = self._kde(good_matches, 0.1)
        p = 1 / (density + 1)
        p[density < 10] = 1e-7
        keep = torch.multinomial(p, num_samples=min(num, len(good_certainty)), replacement=False)
        return good_matches[keep], good_certainty[keep]

    @staticmethod
    def _kde(x, std):
        """utils/kde.py:17-26.  CUDA: `gimb_kde_density` (no n x n distance matrix); CPU tensors: the reference formula."""
        if x.device.type != "cuda":
            return (-torch.cdist(x, x) ** 2 / (2 * std ** 2)).exp().sum(dim=-1)
        x = x.to(torch.float32).contiguous()
        out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().gimb_kde_density(x.data_ptr(), x.shape[0], float(std), out.data_ptr(),
                                                    torch.cuda.current_stream(x.device).cuda_stream))
        return out

    def to_pixel_coordinates(self, matches, H_A, W_A, H_B, W_B):  # dkm.py:652-656
        kA, kB = matches[..., :2], matches[..., 2:]
        kA = torch.stack((W_A / 2 * (kA[..., 0] + 1), H_A / 2 * (kA[..., 1] + 1)), dim=-1)
        kB = torch.stack((W_B / 2 * (kB[..., 0] + 1), H_B / 2 * (kB[..., 1] + 1)), dim=-1)
        return kA, kB

    def forward(self, *a, **k):
        raise NotImplementedError("the training-time forward of RegressionMatcher is not part of the inference path; use match()")


def DKMv3(weights, h, w, symmetric=True, sample_mode="threshold_balanced", **kwargs):
    """networks/dkm/models/model_zoo/DKMv3.py:5 - like the reference, `weights` is not loaded here (its
    `load_state_dict` line is commented out, DKMv3.py:144): callers load the checkpoint themselves (demo.py:355-372)."""
    return RegressionMatcher(h=h, w=w, name="DKMv3", sample_mode=sample_mode, symmetric=symmetric, **kwargs)


# ----------------------------------------------------------------------------- hloc matcher wrapper
def get_padding_size(image, h, w):
    """tools/__init__.py:202-219: pad to the aspect ratio w:h, centred."""
    ow, oh = image.shape[3], image.shape[2]
    ar = w / h
    nw, nh = max(ow, int(oh * ar)), max(oh, int(ow / ar))
    ph, pw = nh - oh, nw - ow
    return ow, oh, pw // 2, pw - pw // 2, ph // 2, ph - ph // 2


class HlocDKM(nn.Module):
    """hloc/matchers/dkm.py:15-154 (the class the reference names `LoFTR` there): pads both images to 3:4, runs
    match() + sample(8192), maps to pixels, removes the padding and keeps in-bounds matches.  The reference's semantic
    masks (.npy files under $GIMRECONSTRUCTION/../segment) are optional here: pass data['mask0'/'mask1'] (bool, True =
    keep) and the images are multiplied by them exactly like dkm.py:63-90."""
    default_conf = {"max_num_matches": None}
    required_inputs = ["image0", "image1"]

    def __init__(self, conf=None, state_dict=None):
        super().__init__()
        self.conf = {**self.default_conf, **(conf or {})}
        self.h, self.w = 672, 896
        self.net = DKMv3(None, self.h, self.w, upsample_preds=True)
        if state_dict is not None:
            sd = dict(state_dict.get("state_dict", state_dict))
            for k in list(sd):
                if k.startswith("model."):
                    sd[k.replace("model.", "", 1)] = sd.pop(k)
            for k in list(sd):
                if "encoder.net.fc" in k:
                    sd.pop(k)
            self.net.load_state_dict(sd)

    @torch.no_grad()
    def forward(self, data):
        # "for consistency with hloc pairs, we refine kpts in image0": the reference swaps the two images
        image0, image1 = data["image1"], data["image0"]
        if "mask1" in data:
            image0 = image0 * data["mask1"].to(image0.dtype)
        if "mask0" in data:
            image1 = image1 * data["mask0"].to(image1.dtype)
        ow0, oh0, pl0, pr0, pt0, pb0 = get_padding_size(image0, self.h, self.w)
        ow1, oh1, pl1, pr1, pt1, pb1 = get_padding_size(image1, self.h, self.w)
        image0 = torch.nn.functional.pad(image0, (pl0, pr0, pt0, pb0))
        image1 = torch.nn.functional.pad(image1, (pl1, pr1, pt1, pb1))
        dense, cert = self.net.match(image0, image1)
        sparse, mconf = self.net.sample(dense, cert, 8192)
        m = mconf > 0
        mconf, sparse = mconf[m], sparse[m]
        h0, w0 = image0.shape[-2:]
        h1, w1 = image1.shape[-2:]
        k0 = torch.stack((w0 * (sparse[:, 0] + 1) / 2, h0 * (sparse[:, 1] + 1) / 2), dim=-1)
        k1 = torch.stack((w1 * (sparse[:, 2] + 1) / 2, h1 * (sparse[:, 3] + 1) / 2), dim=-1)
        k0 = k0 - k0.new_tensor((pl0, pt0))[None]
        k1 = k1 - k1.new_tensor((pl1, pt1))[None]
        keep = (k0[:, 0] > 0) & (k0[:, 1] > 0) & (k1[:, 0] > 0) & (k1[:, 1] > 0)
        keep &= (k0[:, 0] <= ow0 - 1) & (k1[:, 0] <= ow1 - 1) & (k0[:, 1] <= oh0 - 1) & (k1[:, 1] <= oh1 - 1)
        k0, k1, scores = k0[keep], k1[keep], mconf[keep]
        top_k = self.conf["max_num_matches"]
        if top_k is not None and len(scores) > top_k:
            order = torch.argsort(scores, descending=True)[:top_k]
            k0, k1, scores = k0[order], k1[order], scores[order]
        # switch the indices back
        return {"keypoints0": k1, "keypoints1": k0, "scores": scores, "batch_indexes": torch.zeros_like(scores, dtype=torch.long)}
