"""Modules surface of the reference (Modules.py) on torch CUDA tensors.

``Decoder_LSTM`` -- the hot path -- dispatches the whole tf.while_loop of Modules.py:76-119,148-472 (helper, two
ZoneoutLSTMCells, location-sensitive attention, projection, dynamic decode) to the persistent sm_100a kernels through the
C ABI (``decoder.py`` -> ``libmstts_b200.so``), as one ``torch.autograd.Function`` whose backward is the persistent
reverse-time kernel.  There is no CPU path: non-CUDA tensors raise.

``Encoder_Embedding / Encoder_Conv / Encoder_BiLSTM / Decoder_Conv`` are the callers either side of the decoder (SURVEY
8f rank 1): the convolutions on the hand-written tcgen05 GEMM (csrc/conv1d.cu), the zoneout-LSTM sequences and the fused
activation + batch norm + dropout as library kernels of this repository, embedding / losses as torch ops, composed with the
reference's exact semantics (TF batch-norm statistics over padding, dropout/zoneout conventions, sequence-length
handling of stack_bidirectional_dynamic_rnn), with explicit variable dicts instead of TF variable scopes.
"""
from collections import namedtuple

import torch
import torch.nn.functional as F

from . import Hyper_Parameters as hp
from . import _lib
from .decoder import decoder_forward, decoder_backward, fill_mask
from .ZoneoutLSTMCell import ZoneoutLSTMCell  # noqa: F401  (re-exported like the reference module does)
from .Location_Sensitive_Attention import Location_Sensitive_Attention, VARIABLE_KEYS as _ATT_KEYS

# fp32 parity gate (mel L_inf < 1e-3 against the fp32 reference): cuDNN convolutions, forward and backward, stay in fp32
torch.backends.cudnn.allow_tf32 = False

DECODER_KEYS = [key for _, key in _lib.DECODER_WEIGHT_FIELDS]
DECODER_OWN_KEYS = [k for k in DECODER_KEYS if k not in _ATT_KEYS]  # prenet, cells, projection


class Decoder_Output(namedtuple('Decoder_Output', ('linear', 'stop'))):
    pass


class Alignment_History(object):
    """stands in for the TensorArray in AttentionWrapperState.alignment_history: ``stack()`` -> [T, B, Te]"""

    def __init__(self, align_btx):
        self._a = align_btx  # [B, T, Te]

    def stack(self):
        return self._a.transpose(0, 1)


class Decoder_State(namedtuple('Decoder_State', ('time', 'alignment_history'))):
    pass


_WORKSPACE = {}
_WORKSPACE_GEN = {}


def _workspace(dev, nbytes):
    """One cached decoder workspace per device (3.6 GB at B=32, Te=128, L=800 in bf16x3 mode): the saved activations of
    the latest forward live in it until its backward ran, so at most one decoder call may be in flight per device.  Every
    training forward stamps the workspace with a new generation; a backward whose stamp is stale raises instead of
    differentiating through another call's activations.  Returns (workspace, generation)."""
    ws = _WORKSPACE.get(dev)
    if ws is None or ws.numel() < nbytes:
        _WORKSPACE[dev] = ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    gen = _WORKSPACE_GEN.get(dev, 0) + 1
    _WORKSPACE_GEN[dev] = gen
    return ws, gen


class _DecoderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, memory, text_len, mel, mel_len, prenet_mask, zone_mask, n_steps, mode, *weights):
        w = dict(zip(DECODER_KEYS, weights))
        B, Te, D = memory.shape
        m = _lib.MODES[mode]
        nbytes = _lib.lib().mstts_decoder_workspace_bytes(B, Te, mel.shape[1], D, n_steps, m)
        ws, gen = _workspace(memory.device, nbytes)
        lin, stop, align, st = decoder_forward(w, memory, text_len, mel, mel_len, prenet_mask, zone_mask, True, n_steps,
                                               mode, workspace=ws)
        ctx.state, ctx.weights, ctx.ws_gen, ctx.ws_dev = st, w, gen, memory.device
        ctx.mark_non_differentiable(align)
        return lin, stop, align

    @staticmethod
    def backward(ctx, d_linear, d_stop, _d_align):
        if _WORKSPACE_GEN.get(ctx.ws_dev) != ctx.ws_gen:
            raise RuntimeError(
                "Decoder_LSTM: the saved activations of this forward were overwritten by a later training forward on %s "
                "(generation %d, now %s).  One decoder call may be in flight per device: run backward before the next "
                "forward (gradient accumulation = one backward per forward)." % (ctx.ws_dev, ctx.ws_gen, _WORKSPACE_GEN.get(ctx.ws_dev)))
        grads, d_memory = decoder_backward(ctx.state, ctx.weights, d_linear.contiguous(), d_stop.contiguous(),
                                           want_d_memory=ctx.needs_input_grad[0])
        return (d_memory, None, None, None, None, None, None, None) + tuple(grads[k] for k in DECODER_KEYS)


def decoder_mode(B, Te, D):
    """tensor-core (bf16x3) loop where its tiling applies -- any batch size (rows beyond one launch's 32, or 16 for texts
    of 129 .. 256 positions, run as further chunks of the same kernels), texts up to 256 positions -- the fp32 SIMT loop
    otherwise (both meet the 1e-3 gate)"""
    return "bf16x3" if (Te <= 256 and D in (256, 512, 768)) else "fp32"


def Decoder_LSTM(inputs, sequence_length, attention_mechanism, is_training=False, variables=None, masks=None, mode=None,
                 seed=0, max_length=None):
    """Modules.py:76-119.  inputs = Mel [B, L, 80] and sequence_length = Mel_Length [B] (ignored at inference),
    attention_mechanism = Location_Sensitive_Attention (memory, memory_length and the attention variables),
    variables = the decoder's own variables (prenet_*, cell_*, projection/*; short keys of synthetic.TF_VARIABLE_NAMES).
    masks = (prenet_mask [T,2,B,256] u8, zone_mask [T,2,2,B,1024] u8) or None (drawn on device from ``seed``).
    Returns (Decoder_Output(linear [B,T,80], stop [B,T,1]), state) with state.alignment_history.stack() -> [T,B,Te];
    T = max(Mel_Length) + 1 in training, the executed steps at inference.  ``max_length`` = max(Mel_Length) when the caller
    already knows it on the host (the feed dict arrives as host arrays): saves the device read-back, which would stall the host
    until the encoder has drained instead of letting it prepare the decoder launch meanwhile."""
    att = attention_mechanism
    if not isinstance(att, Location_Sensitive_Attention):
        raise TypeError("attention_mechanism must be a Location_Sensitive_Attention")
    if variables is None:
        raise ValueError("Decoder_LSTM needs the decoder variables (there are no implicit TF variable scopes here)")
    w = dict(att.variables)
    w.update({k: variables[k] for k in DECODER_OWN_KEYS})
    memory, text_len = att.memory, att.memory_length
    B, Te, D = memory.shape
    dev = memory.device
    training = bool(is_training)
    # the kernels compile the prenet dropout scale and the zoneout keep factor in (csrc/common.cuh kZoneKeep, 1 / 0.5)
    if hp.Decoder.PreNet.Dropout_Rate != 0.5 or hp.Decoder.LSTM.Zoneout_Rate != 0.1:
        raise ValueError("Decoder_LSTM: the decoder kernels are built for PreNet.Dropout_Rate = 0.5 and LSTM.Zoneout_Rate = 0.1 "
                         "(got %r / %r); rebuild csrc/common.cuh with the new rates" %
                         (hp.Decoder.PreNet.Dropout_Rate, hp.Decoder.LSTM.Zoneout_Rate))
    if training:
        T = (int(max_length) if max_length is not None else int(sequence_length.max().item())) + 1
    else:
        T = hp.Decoder.LSTM.Max_Inference_Length + 1
    if masks is None:
        pm = torch.empty(T, 2, B, 256, device=dev, dtype=torch.uint8)
        fill_mask(pm, 1.0 - hp.Decoder.PreNet.Dropout_Rate, seed * 4 + 1)
        zm = None
        if training:
            zm = torch.empty(T, 2, 2, B, 1024, device=dev, dtype=torch.uint8)
            fill_mask(zm, 1.0 - hp.Decoder.LSTM.Zoneout_Rate, seed * 4 + 2)
    else:
        pm, zm = masks
    if training:
        mode = mode or decoder_mode(B, Te, D)
        lin, stop, align = _DecoderFunction.apply(memory, text_len, inputs, sequence_length, pm, zm, T, mode,
                                                  *[w[k] for k in DECODER_KEYS])
    else:
        # free-running decode: the tcgen05 loop (projection + prenet inside the kernel) for up to 32 rows, else the SIMT kernel
        mode = mode or ("bf16x3" if (B <= 32 and Te <= 128 and D in (256, 512, 768)) else "fp32")
        with torch.no_grad():
            lin, stop, align, _ = decoder_forward(w, memory, text_len, None, None, pm, None, False, T, mode)
    return (Decoder_Output(linear=lin, stop=stop.unsqueeze(-1)),
            Decoder_State(time=lin.shape[1], alignment_history=Alignment_History(align)))


# ------------------------------------------------------------------------------------------------------------------
# callers either side of the decoder (library ops, reference semantics)
# ------------------------------------------------------------------------------------------------------------------
def _dropout(x, rate, training, mask):
    """tf.layers.dropout: keep-prob 1-rate, kept values scaled by 1/(1-rate); identity when not training"""
    if not training:
        return x
    if mask is None:
        mask = torch.floor(torch.rand_like(x) + (1.0 - rate))
    return x / (1.0 - rate) * mask


def _batch_norm(x, v, prefix, training, momentum=0.99, eps=1e-3):
    """tf.layers.batch_normalization on [B,T,C] (axis -1): biased batch statistics over (B,T) *including padding* in
    training, moving statistics otherwise; moving stats are updated in place (the UPDATE_OPS group, MSTTS_SV.py:178)."""
    gamma, beta = v[prefix + '/gamma'], v[prefix + '/beta']
    if training:
        var, mean = torch.var_mean(x, dim=(0, 1), unbiased=False)
        with torch.no_grad():
            v[prefix + '/moving_mean'].mul_(momentum).add_(mean, alpha=1.0 - momentum)
            v[prefix + '/moving_variance'].mul_(momentum).add_(var, alpha=1.0 - momentum)
    else:
        mean, var = v[prefix + '/moving_mean'], v[prefix + '/moving_variance']
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


class _Conv1dFunction(torch.autograd.Function):
    """conv1d 'same' + gradients as bf16x3 tensor-core GEMMs (csrc/conv1d.cu)"""

    @staticmethod
    def _ws(x, B, T, Cin, Cout, k):
        return torch.empty(_lib.lib().mstts_conv1d_workspace_bytes(B, T, Cin, Cout, k), device=x.device, dtype=torch.uint8)

    @staticmethod
    def forward(ctx, x, kernel, bias):
        import ctypes as C
        B, T, Cin = x.shape
        k, _, Cout = kernel.shape
        x, kernel, bias = x.contiguous(), kernel.contiguous(), bias.contiguous()
        y = torch.empty(B, T, Cout, device=x.device)
        ws = _Conv1dFunction._ws(x, B, T, Cin, Cout, k)
        with _lib.on_device(x.device):
            rc = _lib.lib().mstts_conv1d_fwd(_lib.ptr(x), _lib.ptr(kernel), _lib.ptr(bias), B, T, Cin, Cout, k, _lib.ptr(y),
                                             C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr(x.device))
        _lib.check(rc, "mstts_conv1d_fwd")
        ctx.save_for_backward(x, kernel)
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes as C
        x, kernel = ctx.saved_tensors
        B, T, Cin = x.shape
        k, _, Cout = kernel.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dk = torch.empty_like(kernel)
        ws = _Conv1dFunction._ws(x, B, T, Cin, Cout, k)
        with _lib.on_device(x.device):
            rc = _lib.lib().mstts_conv1d_bwd(_lib.ptr(x), _lib.ptr(kernel), _lib.ptr(dy), B, T, Cin, Cout, k, _lib.ptr(dx), _lib.ptr(dk),
                                             C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr(x.device))
        _lib.check(rc, "mstts_conv1d_bwd")
        return dx, dk, dy.sum(dim=(0, 1))


class _ActBnDropoutFunction(torch.autograd.Function):
    """activation -> batch norm -> dropout of a conv layer, fused (csrc/act_bn_dropout.cu)"""

    @staticmethod
    def forward(ctx, x, gamma, beta, moving_mean, moving_var, mask, act, training, rate):
        import ctypes as C
        lib = _lib.lib()
        B, T, Cc = x.shape
        x = x.contiguous()
        y = torch.empty_like(x)
        stats = torch.empty(2, Cc, device=x.device)
        a_saved = torch.empty_like(x) if training else None
        ws = torch.empty(lib.mstts_act_bn_dropout_workspace_bytes(Cc), device=x.device, dtype=torch.uint8)
        keep = 1.0 - rate
        with _lib.on_device(x.device):
            rc = lib.mstts_act_bn_dropout_fwd(_lib.ptr(x), _lib.ptr(gamma.contiguous()), _lib.ptr(beta.contiguous()), _lib.ptr(moving_mean),
                                              _lib.ptr(moving_var), _lib.ptr(mask), B * T, Cc, act, int(training), keep, 0.99, 1e-3,
                                              _lib.ptr(y), _lib.ptr(a_saved), _lib.ptr(stats), C.c_void_p(ws.data_ptr()), ws.numel(),
                                              _lib.stream_ptr(x.device))
        _lib.check(rc, "mstts_act_bn_dropout_fwd")
        if training:
            ctx.save_for_backward(a_saved, stats, gamma, mask)
        ctx.meta = (B, T, Cc, act, keep)
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes as C
        lib = _lib.lib()
        a_saved, stats, gamma, mask = ctx.saved_tensors
        B, T, Cc, act, keep = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        dgamma = torch.empty(Cc, device=dy.device)
        dbeta = torch.empty(Cc, device=dy.device)
        ws = torch.empty(lib.mstts_act_bn_dropout_workspace_bytes(Cc), device=dy.device, dtype=torch.uint8)
        with _lib.on_device(dy.device):
            rc = lib.mstts_act_bn_dropout_bwd(_lib.ptr(dy), _lib.ptr(a_saved), _lib.ptr(stats), _lib.ptr(gamma.contiguous()), _lib.ptr(mask),
                                              B * T, Cc, act, keep, _lib.ptr(dx), _lib.ptr(dgamma), _lib.ptr(dbeta),
                                              C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr(dy.device))
        _lib.check(rc, "mstts_act_bn_dropout_bwd")
        return dx, dgamma, dbeta, None, None, None, None, None, None


def _act_bn_dropout(x, v, prefix, act, training, rate, mask):
    """act (0 relu / 1 tanh) -> tf.layers.batch_normalization -> tf.layers.dropout: one fused kernel pair on CUDA fp32 tensors,
    the library-op composition otherwise (CPU tests, other dtypes)"""
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or v[prefix + '/gamma'].requires_grad or v[prefix + '/beta'].requires_grad)
    if x.is_cuda and x.dtype == torch.float32 and x.shape[-1] % 4 == 0 and (training or not needs_grad):
        # (gradients through the inference-mode normalisation -- moving statistics -- take the library composition below)
        m8 = None
        if training:
            if mask is None:
                m8 = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
                fill_mask(m8, 1.0 - rate, int(torch.randint(0, 2 ** 62, (1,)).item()))
            else:
                m8 = mask.to(torch.uint8).contiguous()
        return _ActBnDropoutFunction.apply(x, v[prefix + '/gamma'], v[prefix + '/beta'], v[prefix + '/moving_mean'],
                                           v[prefix + '/moving_variance'], m8, act, bool(training), rate)
    a = torch.relu(x) if act == 0 else torch.tanh(x)
    a = _batch_norm(a, v, prefix, training)
    return _dropout(a, rate, training, mask)


def _conv1d_same(x, kernel, bias):
    """tf.layers.conv1d(padding='same', stride 1) on [B,T,C] with a TF-layout kernel [k, in, out]: bf16x3 tensor-core GEMMs on
    CUDA fp32 tensors (csrc/conv1d.cu), the library convolution otherwise"""
    k = kernel.shape[0]
    if x.is_cuda and x.dtype == torch.float32 and k % 2 == 1 and k <= 15 and kernel.shape[1] % 8 == 0 and kernel.shape[2] % 8 == 0:
        return _Conv1dFunction.apply(x, kernel, bias)
    y = F.conv1d(x.transpose(1, 2), kernel.permute(2, 1, 0), bias, padding=k // 2)
    return y.transpose(1, 2)


def Encoder_Embedding(inputs, variables):
    """Modules.py:15-23"""
    return F.embedding(inputs.long(), variables['encoder/embedding_variable'])


def Encoder_Conv(inputs, is_training=False, variables=None, masks=None):
    """Modules.py:25-47: conv -> ReLU -> batch norm -> dropout, x3"""
    x = inputs
    for i in range(hp.Encoder.Conv.Nums):
        p = 'encoder/conv_%d' % i
        x = _conv1d_same(x, variables[p + '/conv1d/kernel'], variables[p + '/conv1d/bias'])
        x = _act_bn_dropout(x, variables, p + '/batch_normalization', 0, is_training, hp.Encoder.Conv.Dropout_Rate,
                            None if masks is None else masks[i])
    return x


def _gemm(tA, tB, M, N, K, A, lda, Bm, ldb, Cm, ldc, beta=0.0, precise=0):
    """C[M,N] = op(A) op(B) + beta C on the hand-written tcgen05 kernel (mstts_gemm_f32: bf16x3, fp32 accumulation)"""
    import ctypes as C
    with _lib.on_device(A.device):
        rc = _lib.lib().mstts_gemm_f32(int(tA), int(tB), M, N, K, _lib.ptr(A), lda, 0, _lib.ptr(Bm), ldb, 0, _lib.ptr(Cm), ldc, 0,
                                       float(beta), 1, int(precise), _lib.stream_ptr(A.device))
    _lib.check(rc, "mstts_gemm_f32")


class _DenseFunction(torch.autograd.Function):
    """y[M,N] = x[M,K] w[K,N] (+ bias) and its gradients as three products on the library's own GEMM: the input rows of the
    zoneout-LSTM kernels applied to all steps at once, and tf.layers.dense of the speaker-embedding net"""

    @staticmethod
    def forward(ctx, x, w, bias):
        x, w = x.contiguous(), w.contiguous()
        M, K = x.shape
        N = w.shape[1]
        if bias is not None:
            y = bias.detach().reshape(1, N).expand(M, N).contiguous()
            _gemm(0, 0, M, N, K, x, K, w, N, y, N, beta=1.0)
        else:
            y = torch.empty(M, N, device=x.device)
            _gemm(0, 0, M, N, K, x, K, w, N, y, N)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        M, K = x.shape
        N = w.shape[1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _gemm(0, 1, M, K, N, dy, N, w, N, dx, K)           # dy w^T  (w stored [K, N] = the N x K operand, transposed)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            _gemm(1, 0, K, N, M, x, K, dy, N, dw, N)           # x^T dy
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(dim=0)
        return dx, dw, db


def dense(x, w, bias=None):
    """x[..., K] w[K, N] + bias on CUDA fp32 tensors through the library's GEMM; anything else through torch"""
    if not (x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32):
        y = x @ w
        return y if bias is None else y + bias
    lead = x.shape[:-1]
    y = _DenseFunction.apply(x.reshape(-1, x.shape[-1]), w, bias)
    return y.reshape(*lead, w.shape[1])


_SIDE_STREAM = {}


class side_stream(object):
    """``with side_stream(device, name):`` runs the block on one cached side stream per (device, name), ordered after everything already
    queued on the current stream; ``join(*tensors)`` makes the current stream wait for it and hands the tensors over.  Branches
    of the graph that do not depend on each other (the two directions of the encoder BiLSTM, the frozen speaker-embedding
    net beside the encoder) run concurrently this way -- each is a few-CTA recurrence that leaves most SMs idle.  autograd
    replays every node on the stream its forward ran on, so the reverse passes overlap as well."""

    def __init__(self, device, name='branch'):
        self.dev = torch.device(device)
        self.on = self.dev.type == 'cuda'
        if self.on:
            key = (self.dev.index if self.dev.index is not None else torch.cuda.current_device(), name)
            if key not in _SIDE_STREAM:
                _SIDE_STREAM[key] = torch.cuda.Stream(device=self.dev)
            self.side = _SIDE_STREAM[key]
            self.cur = torch.cuda.current_stream(self.dev)
            self.ctx = torch.cuda.stream(self.side)

    def __enter__(self):
        if self.on:
            self.side.wait_stream(self.cur)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.ctx.__exit__(*exc)
        return False

    def join(self, *tensors):
        if self.on:
            self.cur.wait_stream(self.side)
            for t in tensors:
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(self.cur)
        return tensors[0] if len(tensors) == 1 else tensors


def _reverse_sequence(x, lengths):
    """tf.reverse_sequence over axis 1: the first lengths[b] entries of row b are reversed, the rest stay"""
    T = x.shape[1]
    t = torch.arange(T, device=x.device)[None, :]
    ln = lengths.long()[:, None]
    idx = torch.where(t < ln, ln - 1 - t, t)
    return torch.gather(x, 1, idx[:, :, None].expand_as(x))


class _ZlstmFunction(torch.autograd.Function):
    """the recurrence of a zoneout-LSTM sequence on the GPU (csrc/zlstm.cu): one launch per sequence, forward and reverse"""

    @staticmethod
    def forward(ctx, xk, kh, x_res, lengths, masks, reverse, keep):
        import ctypes as C
        lib = _lib.lib()
        B, T, G = xk.shape
        H = G // 4
        xk, kh = xk.contiguous(), kh.contiguous()
        out = torch.empty(B, T, H, device=xk.device)
        need = xk.requires_grad or kh.requires_grad or (x_res is not None and x_res.requires_grad)
        acts = torch.zeros(B, T, G, device=xk.device) if need else None
        cp = torch.zeros(B, T, H, device=xk.device) if need else None
        hp_ = torch.zeros(B, T, H, device=xk.device) if need else None
        xr = x_res.contiguous() if x_res is not None else None
        with _lib.on_device(xk.device):
            rc = lib.mstts_zlstm_fwd(_lib.ptr(xk), _lib.ptr(kh), _lib.ptr(lengths), _lib.ptr(masks), _lib.ptr(xr), B, T, H, int(reverse),
                                     float(keep), _lib.ptr(out), _lib.ptr(acts), _lib.ptr(cp), _lib.ptr(hp_),
                                     _lib.stream_ptr(xk.device))
        _lib.check(rc, "mstts_zlstm_fwd")
        ctx.save_for_backward(kh, lengths, masks, acts, cp, hp_)
        ctx.meta = (B, T, H, int(reverse), float(keep), x_res is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        import ctypes as C
        kh, lengths, masks, acts, cp, hp_ = ctx.saved_tensors
        B, T, H, reverse, keep, has_res = ctx.meta
        dout = dout.contiguous()
        dxk = torch.empty(B, T, 4 * H, device=dout.device)
        with _lib.on_device(dout.device):
            rc = _lib.lib().mstts_zlstm_bwd(_lib.ptr(dout), _lib.ptr(kh), _lib.ptr(lengths), _lib.ptr(masks), _lib.ptr(acts), _lib.ptr(cp),
                                            B, T, H, reverse, keep, _lib.ptr(dxk),
                                            _lib.stream_ptr(dout.device))
        _lib.check(rc, "mstts_zlstm_bwd")
        dkh = torch.empty(H, 4 * H, device=dout.device)
        _gemm(1, 0, H, 4 * H, B * T, hp_, H, dxk, 4 * H, dkh, 4 * H)  # h_prev^T d xk
        dres = None
        if has_res:
            live = torch.arange(T, device=dout.device)[None, :] < lengths[:, None]
            dres = dout * live[:, :, None]
        return dxk, dkh, dres, None, None, None, None


def zoneout_lstm_sequence(inputs, lengths, kernel, bias, is_training, zoneout_rate, masks=None, residual=False, reverse=False):
    """tf.nn.dynamic_rnn over a ZoneoutLSTMCell with sequence_length: beyond lengths[b] the output is zero and the state
    is carried through.  The input rows of the kernel are applied to all steps in one GEMM; the recurrence (h @ K_h, gates,
    zoneout) is one persistent cluster kernel per sequence on CUDA fp32 tensors with 256 units (csrc/zlstm.cu); other shapes /
    dtypes / devices run the same math step by step with library ops.
    masks: [T, 2, B, H] 0/1 (c, h; indexed by loop step) or None.  residual: tf ResidualWrapper (output = m + input).
    reverse: walk every row from its last valid frame down (reverse_sequence -> rnn -> reverse_sequence)."""
    B, T, In = inputs.shape
    H = kernel.shape[1] // 4
    if inputs.is_cuda and inputs.dtype == torch.float32 and H == 256:
        keep = 1.0 - zoneout_rate
        xk = dense(inputs, kernel[:In], bias)
        m8 = None
        if is_training:
            if masks is None:
                m8 = torch.empty(T, 2, B, H, device=inputs.device, dtype=torch.uint8)
                fill_mask(m8, keep, int(torch.randint(0, 2 ** 62, (1,)).item()))
            else:
                m8 = masks.to(torch.uint8).contiguous()
        out = _ZlstmFunction.apply(xk, kernel[In:], inputs if residual else None, lengths.to(torch.int32).contiguous(), m8, reverse, keep)
        return out, None
    if reverse:
        out, st = zoneout_lstm_sequence(_reverse_sequence(inputs, lengths), lengths, kernel, bias, is_training, zoneout_rate, masks,
                                        residual, False)
        return _reverse_sequence(out, lengths), st
    xk = inputs @ kernel[:In] + bias
    kh = kernel[In:]
    c = inputs.new_zeros(B, H)
    h = inputs.new_zeros(B, H)
    keep = 1.0 - zoneout_rate
    outs = []
    ln = lengths.long()
    for t in range(T):
        g = xk[:, t] + h @ kh
        i, j, f, o = torch.split(g, H, dim=1)
        cn = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
        m = torch.sigmoid(o) * torch.tanh(cn)
        dc, dm = cn - c, m - h
        if is_training:
            if masks is None:
                dc = dc * torch.floor(torch.rand_like(dc) + keep)
                dm = dm * torch.floor(torch.rand_like(dm) + keep)
            else:
                dc, dm = dc * masks[t, 0], dm * masks[t, 1]
        live = (t < ln)[:, None]
        out = m + inputs[:, t] if residual else m
        outs.append(torch.where(live, out, torch.zeros_like(out)))
        c = torch.where(live, keep * dc + c, c)
        h = torch.where(live, keep * dm + h, h)
    return torch.stack(outs, 1), (c, h)


def Encoder_BiLSTM(inputs, lengths, is_training=False, variables=None, masks=None):
    """Modules.py:49-73: stack_bidirectional_dynamic_rnn of ZoneoutLSTMCells (1 layer x 256 units per direction)"""
    x = inputs
    for n in range(hp.Encoder.BiLSTM.Nums):
        p = 'encoder/bilstm/stack_bidirectional_rnn/cell_%d/bidirectional_rnn' % n
        mf = mb = None
        if masks is not None:
            mf, mb = masks[n]
        branch = side_stream(x.device, 'bilstm_bw')
        with branch:  # the reverse direction runs beside the forward one
            bw, _ = zoneout_lstm_sequence(x, lengths, variables[p + '/bw/zoneout_lstm_cell/kernel'],
                                          variables[p + '/bw/zoneout_lstm_cell/bias'], is_training,
                                          hp.Encoder.BiLSTM.Zoneout_Rate, mb, reverse=True)
        fw, _ = zoneout_lstm_sequence(x, lengths, variables[p + '/fw/zoneout_lstm_cell/kernel'],
                                      variables[p + '/fw/zoneout_lstm_cell/bias'], is_training,
                                      hp.Encoder.BiLSTM.Zoneout_Rate, mf)
        bw = branch.join(bw)
        x = torch.cat([fw, bw], dim=-1)
    return x


def Decoder_Conv(inputs, is_training=False, variables=None, masks=None):
    """Modules.py:121-143 (postnet): conv -> tanh -> batch norm -> dropout on all 5 layers incl. the last; the dropout
    rate is read from hp.Encoder.Conv (quirk B-10)"""
    x = inputs
    for i in range(hp.Decoder.Conv.Nums):
        p = 'decoder/conv_%d' % i
        x = _conv1d_same(x, variables[p + '/conv1d/kernel'], variables[p + '/conv1d/bias'])
        x = _act_bn_dropout(x, variables, p + '/batch_normalization', 1, is_training, hp.Encoder.Conv.Dropout_Rate,
                            None if masks is None else masks[i])
    return x
