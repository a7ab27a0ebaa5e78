"""TractOracle-Net scorer (reference: oracles/oracle.py:11-89, transformer_oracle.py:37-118).

``OracleSingleton(checkpoint, device).predict(streamlines) -> np.ndarray[N]`` like the reference;
the resampling to 128 points (dipy ``set_number_of_points``), the ``np.diff`` and the transformer
all run on the device (``ttl_oracle_features`` / ``ttl_oracle_forward*``).  Two precision tiers:
``'fp16'`` (default) is the reference's CUDA arithmetic -- it scores under ``torch.autocast`` fp16
(oracle.py:9,76) -- with the feed-forward blocks on tcgen05 tensor cores (fp16 operands, fp32
accumulators); ``'fp32'`` is the reference's CPU arithmetic on the FP32 pipes, used by the
parity tests.  Every streamline is scored: the reference's own loop drops a trailing partial batch when
N > 4096; we follow ``experiment/oracle_validator.py:40-47``'s chunked semantics.
"""
import ctypes

import numpy as np
import torch

from tracktolearn_b200 import _lib


class TransformerOracleWeights(object):
    """Device copy of a TransformerOracle checkpoint
    ({'hyper_parameters': {...}, 'state_dict': {...}}, transformer_oracle.py:95-118)."""

    def __init__(self, checkpoint, device):
        hp = checkpoint['hyper_parameters']
        if hp.get('name', 'TransformerOracle') != 'TransformerOracle':
            raise ValueError('unsupported oracle model %r' % (hp.get('name'),))
        sd = checkpoint['state_dict']
        self.device = torch.device(device)
        self.n_head = int(hp['n_head'])
        self.n_layers = int(hp['n_layers'])
        self.input_size = int(hp['input_size'])
        self.n_tokens = self.input_size // 3          # 127 directions + CLS
        if self.n_tokens != 128:
            raise _lib.TTLError('oracle kernels are built for 128 tokens (input_size 384), got %d'
                                % self.input_size)

        def dev(t):
            return t.detach().to(self.device, dtype=torch.float32).contiguous()
        self.t = {k: dev(v) for k, v in sd.items() if k != 'pos_encoding.pe'}
        self.t['pe'] = dev(sd['pos_encoding.pe'][:self.n_tokens, 0, :])
        self.d_model = self.t['embedding.0.weight'].shape[0]
        self.d_ff = self.t['bert.layers.0.linear1.weight'].shape[0]
        w = _lib.OracleWeights()
        w.n_layers, w.n_head, w.d_model, w.d_ff, w.n_tokens = (self.n_layers, self.n_head, self.d_model,
                                                               self.d_ff, self.n_tokens)
        w.cls_token = self.t['cls_token'].data_ptr()
        w.emb_w = self.t['embedding.0.weight'].data_ptr()
        w.emb_b = self.t['embedding.0.bias'].data_ptr()
        w.pe = self.t['pe'].data_ptr()
        for i in range(self.n_layers):
            p = 'bert.layers.%d.' % i
            w.in_proj_w[i] = self.t[p + 'self_attn.in_proj_weight'].data_ptr()
            w.in_proj_b[i] = self.t[p + 'self_attn.in_proj_bias'].data_ptr()
            w.out_proj_w[i] = self.t[p + 'self_attn.out_proj.weight'].data_ptr()
            w.out_proj_b[i] = self.t[p + 'self_attn.out_proj.bias'].data_ptr()
            w.lin1_w[i] = self.t[p + 'linear1.weight'].data_ptr()
            w.lin1_b[i] = self.t[p + 'linear1.bias'].data_ptr()
            w.lin2_w[i] = self.t[p + 'linear2.weight'].data_ptr()
            w.lin2_b[i] = self.t[p + 'linear2.bias'].data_ptr()
            w.norm1_w[i] = self.t[p + 'norm1.weight'].data_ptr()
            w.norm1_b[i] = self.t[p + 'norm1.bias'].data_ptr()
            w.norm2_w[i] = self.t[p + 'norm2.weight'].data_ptr()
            w.norm2_b[i] = self.t[p + 'norm2.bias'].data_ptr()
        w.head_w = self.t['head.weight'].data_ptr()
        w.head_b = self.t['head.bias'].data_ptr()
        self.struct = w


class OracleSingleton(object):
    _self = None

    def __new__(cls, *args, **kwargs):
        if cls._self is None:
            cls._self = super().__new__(cls)
        return cls._self

    def __init__(self, checkpoint, device, batch_size=4096, precision='fp16'):
        ck = checkpoint if isinstance(checkpoint, dict) else torch.load(checkpoint, map_location='cpu')
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.TTLError('the oracle runs on a CUDA device only (got %s)' % (self.device,))
        if precision not in ('fp16', 'fp32'):
            raise ValueError("precision must be 'fp16' or 'fp32'")
        self._lib = _lib.load()
        self._destroy_plan()
        self.weights = TransformerOracleWeights(ck, self.device)
        self.batch_size = batch_size
        self.precision = precision
        if precision == 'fp16':
            w = self.weights.struct
            nbytes = self._lib.ttl_oracle_workspace_bytes(ctypes.byref(w))
            if nbytes < 0:
                raise _lib.TTLError('ttl_oracle_workspace_bytes: unsupported oracle shape')
            self._workspace = torch.zeros((nbytes + 1024,), dtype=torch.uint8, device=self.device)
            base = (self._workspace.data_ptr() + 1023) // 1024 * 1024
            plan = ctypes.c_void_p()
            _lib.check(self._lib.ttl_oracle_plan_create(ctypes.byref(plan), ctypes.byref(w), ctypes.c_void_p(base),
                                                        nbytes, _lib.stream_ptr(self.device)),
                       'ttl_oracle_plan_create')
            self._plan = plan

    def _destroy_plan(self):
        plan = getattr(self, '_plan', None)
        if plan is not None:
            self._lib.ttl_oracle_plan_destroy(plan)
        self._plan = None
        self._workspace = None

    def forward_dirs(self, dirs, scores_ptr, n):
        """TransformerOracle.forward on resampled directions [n,127,3] (device) -> scores at ``scores_ptr``."""
        sp = _lib.stream_ptr(self.device)
        if self.precision == 'fp16':
            _lib.check(self._lib.ttl_oracle_forward_tc(self._plan, _lib.ptr(dirs), n, scores_ptr, sp),
                       'ttl_oracle_forward_tc')
        else:
            _lib.check(self._lib.ttl_oracle_forward(ctypes.byref(self.weights.struct), _lib.ptr(dirs), n,
                                                    scores_ptr, sp), 'ttl_oracle_forward')

    @classmethod
    def clear(cls):
        if cls._self is not None and getattr(cls._self, '_lib', None) is not None:
            cls._self._destroy_plan()
        cls._self = None

    # --------------------------------------------------------------------------- device
    def _predict_range(self, points, offsets, start, end):
        """Scores of streamlines [start, end) -> fp32 tensor [end - start] on the device."""
        n = end - start
        scores = torch.empty((max(n, 0),), dtype=torch.float32, device=self.device)
        sp = _lib.stream_ptr(self.device)
        chunk = max(int(self.batch_size), 1)
        dirs = None
        for s0 in range(start, end, chunk):            # bounded scratch: [batch][127][3]
            s1 = min(end, s0 + chunk)
            if dirs is None:
                dirs = torch.empty((min(chunk, n), 127, 3), dtype=torch.float32, device=self.device)
            _lib.check(self._lib.ttl_oracle_features(_lib.ptr(points), ctypes.c_void_p(offsets.data_ptr() + 8 * s0),
                                                     s1 - s0, _lib.ptr(dirs), sp), 'ttl_oracle_features')
            self.forward_dirs(dirs, ctypes.c_void_p(scores.data_ptr() + 4 * (s0 - start)), s1 - s0)
        return scores

    def predict_device(self, points, offsets, distributed=True):
        """points [sum(L),3] fp32 and offsets [N+1] int64 on the device -> scores [N] fp32 (device).

        One process per GPU (``torch.distributed`` initialised, ``distributed=True``): every rank holds
        the same streamlines, scores its contiguous chunk and one ``all_gather_into_tensor`` gives every
        rank all N scores (the reference scores everything on one device,
        oracles/oracle.py:39-89)."""
        n = int(offsets.shape[0]) - 1
        if n <= 0:
            return torch.empty((0,), dtype=torch.float32, device=self.device)
        if not distributed:
            return self._predict_range(points, offsets, 0, n)
        from tracktolearn_b200 import parallel
        return parallel.sharded_scores(n, lambda s, e: self._predict_range(points, offsets, s, e), self.device)

    # --------------------------------------------------------------------------- host API
    def predict(self, streamlines, distributed=True):
        """Reference: oracles/oracle.py:39-89.  ``streamlines``: sequence of [L_i,3] arrays (or an
        object with packed ``data`` / ``offsets``).  Returns float32 numpy [N].  Under
        ``torch.distributed`` the work is shared between the ranks (see ``predict_device``); the oracle
        stopping criterion inside ``step()`` scores rank-local streamlines and passes
        ``distributed=False`` semantics by calling ``forward_dirs`` directly."""
        if hasattr(streamlines, 'offsets') and hasattr(streamlines, 'data'):
            data = np.ascontiguousarray(streamlines.data, dtype=np.float32)
            offsets = np.ascontiguousarray(streamlines.offsets, dtype=np.int64)
        else:
            n = len(streamlines)
            lens = np.fromiter((len(s) for s in streamlines), dtype=np.int64, count=n)
            offsets = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
            data = (np.concatenate([np.asarray(s, dtype=np.float32) for s in streamlines])
                    if n else np.zeros((0, 3), np.float32))
        if len(offsets) <= 1:
            return np.zeros((0,), dtype=np.float32)
        # grow-only pinned staging buffers (a fresh cudaHostAlloc per call costs more than the copy)
        h_pts = self._pinned('pts', data.size, torch.float32)
        h_off = self._pinned('off', offsets.size, torch.int64)
        h_pts.copy_(torch.from_numpy(data.reshape(-1)))
        h_off.copy_(torch.from_numpy(offsets))
        pts = h_pts.to(self.device, non_blocking=True).view(-1, 3)
        off = h_off.to(self.device, non_blocking=True)
        scores = self.predict_device(pts, off, distributed=distributed)
        h_out = self._pinned('scores', scores.numel(), torch.float32)
        h_out.copy_(scores, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return h_out.numpy().copy()

    def _pinned(self, name, numel, dtype):
        cache = self.__dict__.setdefault('_pinned_cache', {})
        buf = cache.get(name)
        if buf is None or buf.numel() < numel or buf.dtype != dtype:
            buf = torch.empty((max(int(numel * 1.25), 1024),), dtype=dtype).pin_memory()
            cache[name] = buf
        return buf[:numel]
