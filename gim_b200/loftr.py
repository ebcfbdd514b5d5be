"""Drop-in `LoFTR` module for the gim_loftr path, backed by libgimb200 (hand-written sm_100a CUDA).

Mirrors the reference's public contract (networks/loftr/loftr.py:14-99):

* `LoFTR(config)` is an `nn.Module` whose `state_dict()` has exactly the reference's keys, so
  `load_state_dict()` accepts the shipped checkpoint (keys may carry the `model.` / `matcher.`
  prefix, loftr.py:93-99) and callers can do `.eval().to(device)` (demo.py:400);
* `forward(data)` mutates `data` in place and adds the same keys with the same dtypes, shapes and
  ordering (`b_ids/i_ids/j_ids/m_bids` int64, `mkpts*` fp32 [M,2] (x, y), `mconf` fp32 [M], ...).

The PyTorch modules below are parameter containers only - their `forward` is never called.  All
compute happens in the CUDA library through the C ABI in include/gimb200.h; there is no fallback.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from .config import get_default_config, lower_config
from .weights import pack_loftr_blob, position_encoding_table


# ----------------------------------------------------------------------------- parameter containers
def _conv(cin, cout, k):
    return nn.Conv2d(cin, cout, kernel_size=k, bias=False)


class _Bottleneck(nn.Module):
    def __init__(self, cin, planes, downsample):
        super().__init__()
        self.conv1, self.bn1 = _conv(cin, planes, 1), nn.BatchNorm2d(planes)
        self.conv2, self.bn2 = _conv(planes, planes, 3), nn.BatchNorm2d(planes)
        self.conv3, self.bn3 = _conv(planes, planes * 4, 1), nn.BatchNorm2d(planes * 4)
        if downsample:
            self.downsample = nn.Sequential(_conv(cin, planes * 4, 1), nn.BatchNorm2d(planes * 4))


class _Trunk(nn.Module):
    """Parameters of the ResNet-50 trunk without max-pool / layer4 (backbone/resnet.py:129-235)."""

    def __init__(self):
        super().__init__()
        self.conv1, self.bn1 = _conv(3, 64, 7), nn.BatchNorm2d(64)
        cin = 64
        for li, (planes, nblk) in enumerate(((64, 3), (128, 4), (256, 6)), start=1):
            blocks = []
            for b in range(nblk):
                blocks.append(_Bottleneck(cin, planes, downsample=(b == 0)))
                cin = planes * 4
            setattr(self, f"layer{li}", nn.Sequential(*blocks))


class _Backbone(nn.Module):
    """Parameters of ResNetFPN_8_2 (backbone/resnet.py:247-289)."""

    def __init__(self, dims):
        super().__init__()
        self.encode = _Trunk()
        self.layer3_outconv = _conv(dims[5], dims[3], 1)
        self.layer2_outconv = _conv(dims[4], dims[3], 1)
        self.layer2_outconv2 = nn.Sequential(_conv(dims[3], dims[3], 3), nn.BatchNorm2d(dims[3]), nn.Identity(),
                                             _conv(dims[3], dims[2], 3))
        self.layer1_outconv = _conv(dims[3], dims[2], 1)
        self.layer1_outconv2 = nn.Sequential(_conv(dims[2], dims[2], 3), nn.BatchNorm2d(dims[2]), nn.Identity(),
                                             _conv(dims[2], dims[1], 3))


class _EncoderLayer(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj = nn.Linear(d, d, bias=False)
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d, bias=False)
        self.merge = nn.Linear(d, d, bias=False)
        self.mlp = nn.Sequential(nn.Linear(2 * d, 2 * d, bias=False), nn.Identity(), nn.Linear(2 * d, d, bias=False))
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)


class _Transformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(cfg["d_model"]) for _ in range(2 * cfg["layer_names"])])


# ----------------------------------------------------------------------------- the module
class StagedU8:
    """One batch whose upload + device-side pre-processing has been enqueued by LoFTR.stage_u8: the device staging buffer,
    the event that marks it ready, its geometry, and references that keep the host tensors alive until the forward ran."""

    def __init__(self, buf, ready, geometry, with_scale, h2d_bytes, keep):
        self.buf, self.ready, self.geometry, self.with_scale, self.h2d_bytes, self._keep = buf, ready, geometry, with_scale, h2d_bytes, keep

    def release(self):
        if self.buf is not None:
            self.buf._gimb_busy = False
        self.buf = None
        self._keep = None


class LoFTR(nn.Module):
    def __init__(self, config=None):
        super().__init__()
        config = lower_config(config) if config is not None else get_default_config()
        self.config = config
        self._validate(config)
        self.backbone = _Backbone(config["resnetfpn"]["block_dims"])
        self.loftr_coarse = _Transformer(config["coarse"])
        self.loftr_fine = _Transformer(config["fine"])
        self.return_conf_matrix = False  # `data['conf_matrix']` is [N,L,S] fp32 (2.9 GB at N=32): opt-in
        self._handle = None
        self._handle_device = None
        self._pe_sizes = set()
        self._workspace = None
        self._staging = None
        self._host_out = None
        self._copy_stream = None
        self._stage_pool = []
        self.last_h2d_bytes = 0
        self.last_d2h_bytes = 0
        for p in self.parameters():
            p.requires_grad_(False)
        if config.get("weight") is not None:
            weights = torch.load(config["weight"], map_location="cpu")
            self.load_state_dict(weights.get("state_dict", weights))

    @staticmethod
    def _validate(c):
        ok = (c["backbone_type"] == "ResNetFPN" and tuple(c["resolution"]) == (8, 2)
              and c["fine_window_size"] == 5 and not c["fine_concat_coarse_feat"]
              and list(c["resnetfpn"]["block_dims"]) == [64, 128, 196, 256, 512, 1024]
              and c["coarse"]["d_model"] == 256 and c["coarse"]["nhead"] == 8 and c["coarse"]["layer_names"] == 4
              and c["coarse"]["attention"] == "linear" and c["fine"]["d_model"] == 128 and c["fine"]["nhead"] == 8
              and c["fine"]["layer_names"] == 1 and c["fine"]["attention"] == "linear"
              and c["match_coarse"]["match_type"] == "dual_softmax")
        if not ok:
            raise NotImplementedError("gim_b200.LoFTR implements the gim_loftr configuration "
                                      "(networks/loftr/config.py defaults) only")

    # -- state handling ------------------------------------------------------------------------
    def load_state_dict(self, state_dict, *args, **kwargs):
        state_dict = dict(state_dict)
        for k in list(state_dict.keys()):  # same prefix stripping as loftr.py:93-99
            for p in ("model.", "matcher."):
                if k.startswith(p):
                    state_dict[k.replace(p, "", 1)] = state_dict.pop(k)
                    break
        self._drop_handle()
        return super().load_state_dict(state_dict, *args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._drop_handle()
        return super()._apply(fn, *args, **kwargs)

    def _drop_handle(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().gimb_loftr_destroy(self._handle)
        self._handle = None
        self._pe_sizes = set()

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("gim_b200.LoFTR is inference-only (the reference's training-time sampling, "
                                      "coarse_matching.py:197-227, is outside the hot path)")
        return super().train(False)

    def _device(self):
        return next(self.parameters()).device

    def _ensure_handle(self):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("gim_b200.LoFTR runs on CUDA (sm_100a) only: call .to('cuda') first; "
                               "there is no CPU path")
        if self._handle is not None and self._handle_device == dev:
            return
        lib = _lib.load()
        self._drop_handle()
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        blob = pack_loftr_blob(sd)
        mc = self.config["match_coarse"]
        cfg = _lib.LoftrCfg(float(mc["thr"]), int(mc["border_rm"]), float(mc["dsmax_temperature"]),
                            int(self.config["fine_window_size"]))
        h = ctypes.c_void_p()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        buf = ctypes.create_string_buffer(blob, len(blob))
        _lib.check(lib.gimb_loftr_create(buf, len(blob), ctypes.byref(cfg), idx, ctypes.byref(h)))
        self._handle, self._handle_device = h, dev

    def _ensure_pe(self, hc, wc):
        if (hc, wc) in self._pe_sizes:
            return
        pe = position_encoding_table(self.config["coarse"]["d_model"], hc, wc)
        _lib.check(_lib.load().gimb_loftr_set_pe(self._handle, hc, wc, pe.data_ptr()))
        self._pe_sizes.add((hc, wc))

    def _ensure_workspace(self, n, h0, w0, h1, w1, dev):
        need = ctypes.c_size_t()
        _lib.check(_lib.load().gimb_loftr_workspace_bytes(self._handle, n, h0, w0, h1, w1, ctypes.byref(need)))
        if self._workspace is None or self._workspace.numel() < need.value or self._workspace.device != dev:
            self._workspace = None
            self._workspace = torch.empty(need.value, dtype=torch.uint8, device=dev)
        return self._workspace

    def profile(self, enabled=True):
        self._ensure_handle()
        _lib.check(_lib.load().gimb_loftr_set_profiling(self._handle, int(enabled)))

    def last_profile(self):
        names = (ctypes.c_char_p * 32)()
        ms = (ctypes.c_float * 32)()
        n = ctypes.c_int()
        _lib.check(_lib.load().gimb_loftr_last_profile(self._handle, names, ms, ctypes.byref(n)))
        return {names[i].decode(): ms[i] for i in range(n.value)}

    def launch_count(self):
        return int(_lib.load().gimb_loftr_launch_count(self._handle)) if self._handle else 0

    def corr_fallbacks(self):
        """Forwards that repeated the coarse matching with the exact correlation sweeps (csrc/corr_sweep.cu)."""
        return int(_lib.load().gimb_loftr_corr_fallbacks(self._handle)) if self._handle else 0

    # -- forward -------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, data, taps=None):
        """Update `data` in place (see networks/loftr/loftr.py:43-91 for the reference contract).

        data: 'image0','image1' (shape only), 'color0','color1' fp32 [N,3,H,W] in [0,1] with H,W % 8 == 0;
              optional 'mask0','mask1' [N,H/8,W/8] and 'scale0','scale1' [N,2].
        CUDA inputs run on the current stream; CPU inputs take the host entry point
        (`gimb_loftr_forward_host`: H2D, forward, D2H) and the outputs come back as CPU tensors.
        """
        self._ensure_handle()
        lib = _lib.load()
        dev = self._device()
        c0, c1 = data["color0"], data["color1"]
        if c0.dim() != 4 or c0.shape[1] != 3 or c1.dim() != 4 or c1.shape[1] != 3 or c0.shape[0] != c1.shape[0]:
            raise RuntimeError(f"color0/color1 must be [N,3,H,W] with equal N, got {tuple(c0.shape)}, {tuple(c1.shape)}")
        if tuple(data["image0"].shape[2:]) != tuple(c0.shape[2:]) or tuple(data["image1"].shape[2:]) != tuple(c1.shape[2:]):
            raise RuntimeError("image0/image1 must have the spatial size of color0/color1")
        n, _, h0, w0 = c0.shape
        h1, w1 = c1.shape[2:]
        if any(v % 8 for v in (h0, w0, h1, w1)):
            raise RuntimeError(f"image sizes must be multiples of 8 (the reference FPN fails otherwise), got "
                               f"{h0}x{w0}, {h1}x{w1}")
        host = c0.device.type == "cpu"
        if not host and c0.device != dev:
            raise RuntimeError(f"inputs on {c0.device} but the matcher is on {dev}")
        has_mask, has_scale = "mask0" in data, "scale0" in data

        def prep(t, dtype, target):
            t = t.to(dtype) if t.dtype != dtype else t
            t = t.contiguous()
            return t if t.device == target else t.to(target)

        tgt = torch.device("cpu") if host else dev
        c0, c1 = prep(c0, torch.float32, tgt), prep(c1, torch.float32, tgt)
        m0 = m1 = s0 = s1 = None
        if has_mask:
            m0, m1 = prep(data["mask0"], torch.uint8, tgt), prep(data["mask1"], torch.uint8, tgt)
            if tuple(m0.shape) != (n, h0 // 8, w0 // 8) or tuple(m1.shape) != (n, h1 // 8, w1 // 8):
                raise RuntimeError("mask0/mask1 must be [N, H/8, W/8]")
        if has_scale:
            s0, s1 = prep(data["scale0"], torch.float32, tgt), prep(data["scale1"], torch.float32, tgt)
            if tuple(s0.shape) != (n, 2) or tuple(s1.shape) != (n, 2):
                raise RuntimeError("scale0/scale1 must be [N, 2]")
        hc0, wc0, hc1, wc1 = h0 // 8, w0 // 8, h1 // 8, w1 // 8
        self._ensure_pe(hc0, wc0)
        self._ensure_pe(hc1, wc1)

        with torch.cuda.device(dev):
            ws = self._ensure_workspace(n, h0, w0, h1, w1, dev)
            cap = n * min(hc0 * wc0, hc1 * wc1)
            outs = self._alloc_outputs(cap, dev)
            o = self._out_struct(cap, outs)
            tp = None
            tap_tensors = {}
            want_conf = self.return_conf_matrix or bool(data.get("return_conf_matrix", False))
            if taps or want_conf:
                tap_tensors = self._alloc_taps(set(taps or ()) | ({"conf_matrix"} if want_conf else set()),
                                               n, h0, w0, h1, w1, cap, dev)
                tp = _lib.LoftrTaps(**{k: v.data_ptr() for k, v in tap_tensors.items()})
            m_out = ctypes.c_int64()
            stream = torch.cuda.current_stream(dev).cuda_stream
            ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
            if not host:
                _lib.check(lib.gimb_loftr_forward(self._handle, ptr(c0), ptr(c1), ptr(m0), ptr(m1), ptr(s0), ptr(s1),
                                                  n, h0, w0, h1, w1, ws.data_ptr(), ws.numel(), ctypes.byref(o),
                                                  ctypes.byref(tp) if tp else None, ctypes.byref(m_out), stream))
                M = m_out.value
                res = {k: v[:M] for k, v in outs.items()}
            else:
                if tp is not None:
                    raise RuntimeError("taps / conf_matrix need CUDA inputs")
                need = ctypes.c_size_t()
                _lib.check(lib.gimb_loftr_host_staging_bytes(n, h0, w0, h1, w1, int(has_mask), int(has_scale),
                                                             ctypes.byref(need)))
                if self._staging is None or self._staging.numel() < need.value or self._staging.device != dev:
                    self._staging = torch.empty(need.value, dtype=torch.uint8, device=dev)
                # persistent pinned staging for the D2H copies (pinned allocation per call is slow, and slower still
                # when several ranks allocate concurrently); the caller gets private copies of rows [0, M)
                if self._host_out is None or self._host_out["b_ids"].shape[0] < cap:
                    self._host_out = self._alloc_outputs(cap, torch.device("cpu"), pin=True)
                houts = self._host_out
                ho = self._out_struct(houts["b_ids"].shape[0], houts)
                up, down = ctypes.c_uint64(), ctypes.c_uint64()
                _lib.check(lib.gimb_loftr_forward_host(self._handle, ptr(c0), ptr(c1), ptr(m0), ptr(m1), ptr(s0),
                                                       ptr(s1), n, h0, w0, h1, w1, self._staging.data_ptr(),
                                                       self._staging.numel(), ws.data_ptr(), ws.numel(),
                                                       ctypes.byref(o), ctypes.byref(ho), ctypes.byref(m_out),
                                                       ctypes.byref(up), ctypes.byref(down), stream))
                M = m_out.value
                self.last_h2d_bytes, self.last_d2h_bytes = up.value, down.value
                res = {k: v[:M].clone() for k, v in houts.items()}

        data.update({
            "bs": n, "hw0_i": data["image0"].shape[2:], "hw1_i": data["image1"].shape[2:],
            "hw0_c": torch.Size((hc0, wc0)), "hw1_c": torch.Size((hc1, wc1)),
            "hw0_f": torch.Size((h0 // 2, w0 // 2)), "hw1_f": torch.Size((h1 // 2, w1 // 2)),
            "b_ids": res["b_ids"], "i_ids": res["i_ids"], "j_ids": res["j_ids"],
            "gt_mask": res["mconf"] == 0, "m_bids": res["b_ids"],
            "mkpts0_c": res["mkpts0_c"], "mkpts1_c": res["mkpts1_c"], "mconf": res["mconf"],
            "W": int(self.config["fine_window_size"]), "expec_f": res["expec_f"],
            "mkpts0_f": res["mkpts0_f"], "mkpts1_f": res["mkpts1_f"],
        })
        if "conf_matrix" in tap_tensors:
            data["conf_matrix"] = tap_tensors["conf_matrix"]
        if taps:
            data["_taps"] = {k: (v[:M] if k.startswith("fine_win") else v) for k, v in tap_tensors.items()}

    def _u8_geometry(self, data):
        u0, u1 = data["color0_u8"], data["color1_u8"]
        if u0.dtype != torch.uint8 or u1.dtype != torch.uint8 or u0.dim() != 4 or u1.dim() != 4 or u0.shape[3] != 3 or \
                u1.shape[3] != 3 or u0.shape[0] != u1.shape[0] or u0.device.type != "cpu" or u1.device.type != "cpu":
            raise RuntimeError("color0_u8 / color1_u8 must be CPU uint8 tensors [N, h, w, 3] with equal N")
        u0, u1 = u0.contiguous(), u1.contiguous()
        n, ih0, iw0, _ = u0.shape
        _, ih1, iw1, _ = u1.shape
        h0, w0 = data.get("pad0", (ih0, iw0))
        h1, w1 = data.get("pad1", (ih1, iw1))
        if any(v % 8 for v in (h0, w0, h1, w1)):
            raise RuntimeError(f"padded image sizes must be multiples of 8, got {h0}x{w0}, {h1}x{w1}")
        s0 = s1 = None
        if "scale0" in data:
            s0, s1 = data["scale0"].to(torch.float32).contiguous().cpu(), data["scale1"].to(torch.float32).contiguous().cpu()
        return u0, u1, s0, s1, (n, ih0, iw0, ih1, iw1, h0, w0, h1, w1)

    @torch.no_grad()
    def stage_u8(self, data, stream=None):
        """First half of forward_u8 (`gimb_loftr_stage_host_u8`): enqueue the upload and the device-side pre-processing of one
        batch on `stream` (default: a copy stream owned by the model) and return a StagedU8 WITHOUT waiting, so that the
        upload of batch i+1 overlaps the forward of batch i.  Pass the result as `data['staged']` to forward_u8.  The uint8
        host tensors should be pinned (a pageable copy is synchronous) and must not change until the forward ran."""
        self._ensure_handle()
        lib = _lib.load()
        dev = self._device()
        u0, u1, s0, s1, geo = self._u8_geometry(data)
        n, ih0, iw0, ih1, iw1, h0, w0, h1, w1 = geo
        with torch.cuda.device(dev):
            need = ctypes.c_size_t()
            _lib.check(lib.gimb_loftr_host_u8_staging_bytes(n, ih0, iw0, ih1, iw1, h0, w0, h1, w1, int(s0 is not None), ctypes.byref(need)))
            cur = torch.cuda.current_stream(dev)
            if stream is None:
                if self._copy_stream is None or self._copy_stream.device != dev:
                    self._copy_stream = torch.cuda.Stream(dev)
                stream = self._copy_stream
            # staging buffers are kept by the model and recycled once their forward ran (allocating ~300 MB per call through
            # the caching allocator with a cross-stream dependency costs a cudaMalloc per step)
            buf = None
            for cand in self._stage_pool:
                if cand.device == dev and cand.numel() >= need.value and not getattr(cand, "_gimb_busy", False):
                    buf = cand
                    break
            if buf is None:
                self._stage_pool = [c for c in self._stage_pool if getattr(c, "_gimb_busy", False)][-3:]
                buf = torch.empty(need.value, dtype=torch.uint8, device=dev)
                self._stage_pool.append(buf)
                if stream != cur:
                    buf.record_stream(stream)
            buf._gimb_busy = True
            if stream != cur:
                stream.wait_stream(cur)      # the previous consumer of this buffer ran on the compute stream
            up = ctypes.c_uint64()
            ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
            _lib.check(lib.gimb_loftr_stage_host_u8(self._handle, u0.data_ptr(), ih0, iw0, u1.data_ptr(), ih1, iw1, ptr(s0), ptr(s1),
                                                    n, h0, w0, h1, w1, buf.data_ptr(), buf.numel(), ctypes.byref(up), stream.cuda_stream))
            ready = torch.cuda.Event()
            ready.record(stream)
        return StagedU8(buf, ready, geo, s0 is not None, up.value, (u0, u1, s0, s1))

    @torch.no_grad()
    def forward_u8(self, data):
        """GPU pre-processing entry (`gimb_loftr_forward_host_u8`): `data['color0_u8'|'color1_u8']` are
        CPU uint8 RGB tensors [N, h, w, 3] as cv2 delivers them (after the loader's cv2.resize, datasets/utils.py:108);
        the float conversion (/255), HWC -> CHW, the zero padding to `data['pad0'|'pad1']` = (H, W) (default: the image
        size, which must then be a multiple of 8) and the 1/8 padding masks happen on the device.  Optional
        `scale0/scale1` [N, 2].  Results come back as CPU tensors under the same keys as forward().
        With `data['staged']` = the result of stage_u8() the upload has already been enqueued (possibly on another stream)
        and only the forward + read-back run here."""
        self._ensure_handle()
        lib = _lib.load()
        dev = self._device()
        staged = data.get("staged")
        if staged is None:
            staged = self.stage_u8(data, stream=torch.cuda.current_stream(dev))
        elif staged.buf is None:
            raise RuntimeError("forward_u8: this staged batch has already been consumed")
        n, ih0, iw0, ih1, iw1, h0, w0, h1, w1 = staged.geometry
        hc0, wc0, hc1, wc1 = h0 // 8, w0 // 8, h1 // 8, w1 // 8
        self._ensure_pe(hc0, wc0)
        self._ensure_pe(hc1, wc1)
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(staged.ready)
            ws = self._ensure_workspace(n, h0, w0, h1, w1, dev)
            cap = n * min(hc0 * wc0, hc1 * wc1)
            outs = self._alloc_outputs(cap, dev)
            o = self._out_struct(cap, outs)
            # persistent pinned read-back buffers (pinning ~10 MB per call costs tens of ms); the results are cloned out of them
            if self._host_out is None or self._host_out["b_ids"].shape[0] < cap:
                self._host_out = self._alloc_outputs(cap, torch.device("cpu"), pin=True)
            houts = self._host_out
            ho = self._out_struct(houts["b_ids"].shape[0], houts)
            down, m_out = ctypes.c_uint64(), ctypes.c_int64()
            _lib.check(lib.gimb_loftr_forward_staged_u8(self._handle, ih0, iw0, ih1, iw1, int(staged.with_scale), n, h0, w0, h1, w1,
                                                        staged.buf.data_ptr(), staged.buf.numel(), ws.data_ptr(), ws.numel(),
                                                        ctypes.byref(o), ctypes.byref(ho), ctypes.byref(m_out), ctypes.byref(down),
                                                        cur.cuda_stream))
            M = m_out.value
            self.last_h2d_bytes, self.last_d2h_bytes = staged.h2d_bytes, down.value
            res = {k: v[:M].clone() for k, v in houts.items()}
        staged.release()
        data.update({
            "bs": n, "hw0_i": torch.Size((h0, w0)), "hw1_i": torch.Size((h1, w1)),
            "hw0_c": torch.Size((hc0, wc0)), "hw1_c": torch.Size((hc1, wc1)),
            "hw0_f": torch.Size((h0 // 2, w0 // 2)), "hw1_f": torch.Size((h1 // 2, w1 // 2)),
            "b_ids": res["b_ids"], "i_ids": res["i_ids"], "j_ids": res["j_ids"],
            "gt_mask": res["mconf"] == 0, "m_bids": res["b_ids"],
            "mkpts0_c": res["mkpts0_c"], "mkpts1_c": res["mkpts1_c"], "mconf": res["mconf"],
            "W": int(self.config["fine_window_size"]), "expec_f": res["expec_f"],
            "mkpts0_f": res["mkpts0_f"], "mkpts1_f": res["mkpts1_f"],
        })

    @staticmethod
    def _alloc_outputs(cap, dev, pin=False):
        kw = dict(device=dev)
        if pin:
            kw["pin_memory"] = True
        e = torch.empty
        return {
            "b_ids": e(cap, dtype=torch.int64, **kw), "i_ids": e(cap, dtype=torch.int64, **kw),
            "j_ids": e(cap, dtype=torch.int64, **kw), "mconf": e(cap, dtype=torch.float32, **kw),
            "mkpts0_c": e(cap, 2, dtype=torch.float32, **kw), "mkpts1_c": e(cap, 2, dtype=torch.float32, **kw),
            "mkpts0_f": e(cap, 2, dtype=torch.float32, **kw), "mkpts1_f": e(cap, 2, dtype=torch.float32, **kw),
            "expec_f": e(cap, 3, dtype=torch.float32, **kw),
        }

    @staticmethod
    def _out_struct(cap, t):
        return _lib.LoftrOut(cap, *(t[k].data_ptr() for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_c",
                                                               "mkpts1_c", "mkpts0_f", "mkpts1_f", "expec_f")))

    @staticmethod
    def _alloc_taps(names, n, h0, w0, h1, w1, cap, dev):
        L, S = (h0 // 8) * (w0 // 8), (h1 // 8) * (w1 // 8)
        shapes = {
            "feat_c_backbone0": (n, h0 // 8, w0 // 8, 256), "feat_c_backbone1": (n, h1 // 8, w1 // 8, 256),
            "feat_f0": (n, h0 // 2, w0 // 2, 128), "feat_f1": (n, h1 // 2, w1 // 2, 128),
            "feat_c0": (n, L, 256), "feat_c1": (n, S, 256),
            "fine_win0": (cap, 25, 128), "fine_win1": (cap, 25, 128), "conf_matrix": (n, L, S),
        }
        bad = names - set(shapes)
        if bad:
            raise KeyError(f"unknown taps {sorted(bad)}")
        return {k: torch.empty(shapes[k], dtype=torch.float32, device=dev) for k in names}
