"""Drop-in `warp_learn.models.G_Resnet` (the ICN generator) on hand-written sm_100a kernels.

Contract kept (reference: warp_learn/models.py:190-208, run_test.py:74-78, trajectory_inference.py:182,391):
  * `G_Resnet(input_nc, output_nc=3, num_downs=2, n_res=3, ngf=64, norm='inst', nl_layer='relu')` is an nn.Module;
    `.to()`, `.eval()`, `.load_state_dict(sd, strict=True)` with the reference's 40 keys
    (`enc_content.model.*`, `dec.model.*`, LayerNorm `norm.gamma/beta` registered before `conv.weight/bias`) in the
    reference's order, `.state_dict()` round trip;
  * `forward(image (B,input_nc,H,W) fp32) -> (B,output_nc,H,W)` tanh-bounded fp32 on the module's device;
    `decode(content)`; `enc_content(image)` and `dec(content)` are callable like the reference's sub-modules.
Only the configuration the reference ships (norm='inst', nl_layer='relu', n_res >= 1, num_downs >= 1) has kernels; anything
else raises NotImplementedError.  There is no CPU path.  The input packing `get_icn_inputs` (warp_learn/models.py:17-56) lives in
`frame_ops.py` (`get_icn_inputs`, `get_icn_inputs_batch`) and is what the import shim serves.
"""
import math

import torch
import torch.nn as nn

from .icn_engine import IcnEngine


class _Conv(nn.Module):
    """Parameter holder with nn.Conv2d's names, shapes and default init range."""

    def __init__(self, cin, cout, k):
        super().__init__()
        bound = 1.0 / math.sqrt(cin * k * k)
        self.weight = nn.Parameter((torch.rand(cout, cin, k, k) * 2 - 1) * bound)
        self.bias = nn.Parameter((torch.rand(cout) * 2 - 1) * bound)


class _LN(nn.Module):
    """gamma ~ U(0,1), beta = 0 (warp_learn/models.py:22-24)."""

    def __init__(self, c):
        super().__init__()
        self.gamma = nn.Parameter(torch.rand(c))
        self.beta = nn.Parameter(torch.zeros(c))


class _Block(nn.Module):
    """One Conv2dBlock's parameters: optional LayerNorm first, then the convolution (registration order of models.py:51-81)."""

    def __init__(self, cin, cout, k, ln=False):
        super().__init__()
        self.norm = _LN(cout) if ln else None
        self.conv = _Conv(cin, cout, k)


class _Seq(nn.Module):
    """Container whose child names are the reference's nn.Sequential indices."""

    def __init__(self, fn=None):
        super().__init__()
        self._fn = fn

    def forward(self, x):
        if self._fn is None:
            raise NotImplementedError("this container only holds parameters; call G_Resnet.forward / enc_content / dec")
        return self._fn(x)


class G_Resnet(nn.Module):
    def __init__(self, input_nc, output_nc=3, num_downs=2, n_res=3, ngf=64, norm='inst', nl_layer='relu', dtype: str = "fp16",
                 impl: str = "auto"):
        """dtype: 'fp16' (tcgen05 product path), 'bf16' (same kernels on bf16), 'fp32' (verification build on CUDA cores)."""
        super().__init__()
        if norm != 'inst' or nl_layer != 'relu' or num_downs < 1 or n_res < 1:
            raise NotImplementedError("the B200 path implements the shipped ICN configuration only (norm='inst', nl_layer='relu'; "
                                      "run_test.py:75)")
        self.input_nc, self.output_nc, self.n_down, self.n_res, self.ngf = input_nc, output_nc, num_downs, n_res, ngf
        self.blocks = {}

        def add(scope, path, name, blk):
            scope.add_module(name, blk)
            self.blocks[path] = blk

        def res_blocks(prefix, dim):
            outer = _Seq()
            inner_seq = _Seq()
            outer.add_module("model", inner_seq)
            for r in range(n_res):
                rb = _Seq()
                rb_seq = _Seq()
                rb.add_module("model", rb_seq)
                inner_seq.add_module(str(r), rb)
                for j in range(2):
                    add(rb_seq, f"{prefix}.model.{r}.model.{j}", str(j), _Block(dim, dim, 3))
            return outer

        # ContentEncoder (models.py:128-149)
        self.enc_content = _Seq(self._enc_api)
        enc = _Seq()
        self.enc_content.add_module("model", enc)
        add(enc, "enc_content.model.0", "0", _Block(input_nc, ngf, 7))
        dim = ngf
        for i in range(num_downs):
            add(enc, f"enc_content.model.{1 + i}", str(1 + i), _Block(dim, 2 * dim, 4))
            dim *= 2
        enc.add_module(str(1 + num_downs), res_blocks(f"enc_content.model.{1 + num_downs}", dim))
        # Decoder (models.py:164-188); the Upsample modules at odd indices have no parameters
        self.dec = _Seq(self._dec_api)
        dec = _Seq()
        self.dec.add_module("model", dec)
        dec.add_module("0", res_blocks("dec.model.0", dim))
        for i in range(num_downs):
            add(dec, f"dec.model.{2 + 2 * i}", str(2 + 2 * i), _Block(dim, dim // 2, 5, ln=True))
            dim //= 2
        add(dec, f"dec.model.{1 + 2 * num_downs}", str(1 + 2 * num_downs), _Block(dim, output_nc, 7))
        self._engines = {}
        self._dtype, self._impl = dtype, impl

    # ------------------------------------------------------------------ engine access
    def engine(self) -> IcnEngine:
        key = (self._dtype, self._impl)
        if key not in self._engines:
            self._engines[key] = IcnEngine(self, self._dtype, self._impl)
        eng = self._engines[key]
        eng.prepare_weights()
        return eng

    def set_compute(self, dtype="fp16", impl="auto"):
        self._dtype, self._impl = dtype, impl
        return self

    def _dev_ctx(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            from .._lib import FusgError
            raise FusgError("G_Resnet (B200): parameters are on %s; this path has no CPU fallback -- call .to('cuda')" % dev)
        return torch.cuda.device(dev)

    @staticmethod
    def _check(image):
        if image.dim() != 4 or image.shape[-1] % 4 or image.shape[-2] % 4 or min(image.shape[-2:]) < 16:
            raise NotImplementedError("G_Resnet (B200): input must be (B,C,H,W) with H, W multiples of 4 and >= 16")

    # ------------------------------------------------------------------ reference API
    def _enc_api(self, image):
        self._check(image)
        with torch.no_grad(), self._dev_ctx():
            e = self.engine()
            c = e.encode(image)
            out = e.content_to_nchw(c)
            out._fusg_content = (c, out._version)
            return out

    def _content(self, e, content):
        tag = getattr(content, "_fusg_content", None)
        if tag is not None and tag[1] == content._version:
            return tag[0]
        return e.to_padded(content, 1)

    def _dec_api(self, content):
        with torch.no_grad(), self._dev_ctx():
            e = self.engine()
            return e.decode(self._content(e, content))

    def decode(self, content):
        return self.dec(content)

    def forward(self, image):
        self._check(image)
        with torch.no_grad(), self._dev_ctx():
            return self.engine().forward(image)
