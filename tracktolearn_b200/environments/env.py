"""BaseEnv: subject data in HBM + tracking parameters (reference: environments/env.py).

Same constructor contract as the reference (``cls(subject_data, split_id, env_dto)``,
``from_files``), same public attributes (``seeds``, ``affine_vox2rasmm``, ``step_size_mm``,
``max_nb_steps``, ``target_sh_order``, ``tracking_mask``, ``reference``).  What differs is
where the work happens: the volumes live on the device in the layouts the CUDA step kernels
want and every per-step computation is a kernel call through the C ABI (``_lib``).
"""
import ctypes

import numpy as np
import torch
from scipy.ndimage import spline_filter

from tracktolearn_b200 import _lib
from tracktolearn_b200.datasets.utils import (MRIDataVolume, convert_length_mm2vox,
                                              get_sh_order_and_fullness)
from tracktolearn_b200.environments.utils import random_seeds_from_mask


class BaseEnv(object):
    """Abstract tracking environment (reference: environments/env.py:40-141)."""

    CHANNEL_ALIGN = 4  # floats: every voxel is a whole number of float4

    def __init__(self, subject_data, split_id, env_dto):
        if type(subject_data) is str:
            raise NotImplementedError(
                'HDF5 datasets (env.py:85-94) are outside the hot path this package covers; '
                'pass (input_volume, tracking_mask, seeding_mask, peaks, reference)')
        self.subject_data = subject_data
        self.split = split_id
        self._lib = _lib.load()

        self._state_size = None

        # Tracking parameters (env.py:108-138)
        self.n_dirs = env_dto['n_dirs']
        self.theta = env_dto['theta']
        self.npv = env_dto['npv']
        self.binary_stopping_threshold = env_dto['binary_stopping_threshold']
        self.step_size_mm = env_dto['step_size']
        self.min_length_mm = env_dto['min_length']
        self.max_length_mm = env_dto['max_length']
        self.oracle_checkpoint = env_dto.get('oracle_checkpoint')
        self.oracle_stopping_criterion = env_dto.get('oracle_stopping_criterion', False)
        self.scoring_data = env_dto.get('scoring_data')
        self.compute_reward = env_dto['compute_reward']
        self.alignment_weighting = env_dto.get('alignment_weighting', 0.0)
        self.oracle_bonus = env_dto.get('oracle_bonus', 0.0)
        self.rng = env_dto['rng']
        self.device = torch.device(env_dto['device'])
        if self.device.type != 'cuda':
            raise _lib.TTLError('tracktolearn_b200 environments run on a CUDA device only '
                                '(got %s); there is no CPU fallback' % (self.device,))
        self.target_sh_order = env_dto.get('target_sh_order')
        # Build the state rows of streamlines that stop this step, like the reference does
        # (tracking_env.py:214-215).  Trackers that never read them switch this off.
        self.state_of_stopped = env_dto.get('state_of_stopped', True)
        self._float64_directions = False   # TrackingEnvironment; Noisy env flips it
        # TractOracle-Net inside the step (stopping_criteria.py:85-154, oracle_reward.py:10-93)
        self._oracle = None
        if self.oracle_checkpoint and (self.oracle_stopping_criterion
                                       or (self.compute_reward and self.oracle_bonus > 0)):
            from tracktolearn_b200.oracles.oracle import OracleSingleton
            # 'fp16' = the reference's CUDA arithmetic (autocast), 'fp32' = its CPU arithmetic
            self._oracle = OracleSingleton(self.oracle_checkpoint, self.device,
                                           precision=env_dto.get('oracle_precision', 'fp16'))

        self._uploaded = False
        self.load_subject()

    # ------------------------------------------------------------------ subject / parameters
    def load_subject(self):
        """Reference: environments/env.py:143-282."""
        (input_volume, tracking_mask, seeding_mask, peaks, reference) = self.subject_data

        self.affine_vox2rasmm = input_volume.affine_vox2rasmm
        self.affine_rasmm2vox = np.linalg.inv(self.affine_vox2rasmm)
        self.reference = reference

        if self.target_sh_order is None:
            sh_order, _ = get_sh_order_and_fullness(input_volume.shape[-1])
            self.target_sh_order = sh_order

        self.tracking_mask = tracking_mask
        self.peaks = peaks
        sd = seeding_mask.data
        if isinstance(sd, torch.Tensor):
            sd = sd.cpu().numpy()
        self.seeding_data = np.asarray(sd).astype(np.uint8)

        if not self._uploaded:
            self._upload_volumes(input_volume, tracking_mask, peaks)
            self._uploaded = True

        self.step_size = convert_length_mm2vox(self.step_size_mm, self.affine_vox2rasmm)
        self.min_length = self.min_length_mm
        self.max_length = self.max_length_mm
        self.max_nb_steps = int(self.max_length / self.step_size_mm)
        self.min_nb_steps = int(self.min_length / self.step_size_mm)
        self.add_neighborhood_vox = convert_length_mm2vox(self.step_size_mm, self.affine_vox2rasmm)
        r = np.float32(self.add_neighborhood_vox)
        eye = np.eye(3, dtype=np.float32)
        self.neighborhood_directions = torch.from_numpy(
            np.concatenate((np.zeros((1, 3), np.float32), eye * r, -eye * r))).to(self.device)

        # Tracking seeds (env.py:216-219), global numpy RNG like dipy
        self.seeds = random_seeds_from_mask(self.seeding_data, self.npv)

        self._params = _lib.Params(
            step_vox=float(self.step_size), mask_threshold=float(self.binary_stopping_threshold),
            alignment_weighting=float(self.alignment_weighting),
            theta_rad=float(np.float32(np.deg2rad(self.theta))), max_nb_steps=int(self.max_nb_steps),
            n_dirs=int(self.n_dirs), dir_f64=int(self._float64_directions),
            compute_reward=int(bool(self.compute_reward)), state_stopped=int(bool(self.state_of_stopped)))
        self._batch = None   # buffers depend on max_nb_steps: reallocate lazily

    def _upload_volumes(self, input_volume, tracking_mask, peaks):
        lib = self._lib
        data = input_volume.data
        if not isinstance(data, torch.Tensor):
            data = torch.as_tensor(np.asarray(data))
        if data.dim() != 4:
            raise ValueError('input volume must be [X,Y,Z,C]')
        X, Y, Z, C = data.shape
        CP = (C + self.CHANNEL_ALIGN - 1) // self.CHANNEL_ALIGN * self.CHANNEL_ALIGN
        raw = data.to(self.device, dtype=torch.float32).contiguous()
        # largest coefficient magnitude (NaN-aware): fp16 operand rows hold trilinear combinations of these
        # values, so this one number decides whether the fp16 tier can represent the states of this volume
        self._sh_absmax = float(torch.nan_to_num(raw, nan=float('inf')).abs().max().item()) if raw.numel() else 0.0
        self._sh = torch.empty((X, Y, Z, CP), dtype=torch.float32, device=self.device)
        _lib.check(lib.ttl_pad_channels(_lib.ptr(raw), _lib.ptr(self._sh), X * Y * Z, C, CP,
                                        _lib.stream_ptr(self.device)), 'ttl_pad_channels')
        torch.cuda.current_stream(self.device).synchronize()
        del raw
        self._n_coefs = C
        # stopping_criteria.py:58-59: cubic B-spline coefficients of the mask, float64
        mask_data = tracking_mask.data
        if isinstance(mask_data, torch.Tensor):
            mask_data = mask_data.cpu().numpy()
        mask_data = np.asarray(mask_data).astype(np.uint8)
        coef = spline_filter(np.ascontiguousarray(mask_data, dtype=float), order=3)
        self._mask_coef = torch.from_numpy(coef).to(self.device)
        self._peaks = None
        if peaks is not None and self.compute_reward:
            pk = peaks.data
            if not isinstance(pk, torch.Tensor):
                pk = torch.from_numpy(np.ascontiguousarray(np.asarray(pk, dtype=np.float32)))
            self._peaks = pk.to(self.device, dtype=torch.float32).contiguous()
        pshape = tuple(self._peaks.shape[:3]) if self._peaks is not None else (0, 0, 0)
        self._volume = _lib.Volume(
            sh=self._sh.data_ptr(), X=X, Y=Y, Z=Z, C=C, CP=CP,
            mask_coef=self._mask_coef.data_ptr(), MX=coef.shape[0], MY=coef.shape[1], MZ=coef.shape[2],
            peaks=self._peaks.data_ptr() if self._peaks is not None else None,
            PX=pshape[0], PY=pshape[1], PZ=pshape[2])

    @property
    def data_volume(self):
        """[X,Y,Z,C] fp32 view of the device volume (reference attribute, env.py:179-180)."""
        return self._sh[..., :self._n_coefs]

    # ------------------------------------------------------------------------- constructors
    @classmethod
    def from_dataset(cls, env_dto, split):
        raise NotImplementedError('HDF5 datasets are out of scope (SURVEY.md section 2)')

    @classmethod
    def from_files(cls, env_dto):
        """Reference: environments/env.py:311-347."""
        from tracktolearn_b200.datasets.files import load_files
        (input_volume, peaks_volume, tracking_mask, seeding_mask) = load_files(
            env_dto['in_odf'], env_dto['in_seed'], env_dto['in_mask'], env_dto['sh_basis'],
            env_dto['target_sh_order'], compute_peaks=bool(env_dto.get('compute_reward')),
            device=env_dto.get('device', 'cuda:0'))
        subj_files = (input_volume, tracking_mask, seeding_mask, peaks_volume, env_dto['reference'])
        return cls(subj_files, 'testing', env_dto)

    # ---------------------------------------------------------------------------- accessors
    def get_state_size(self):
        """Reference: env.py:451-463."""
        self._state_size = 7 * self._n_coefs + 3 * self.n_dirs
        return self._state_size

    def get_action_size(self):
        return 3

    def get_target_sh_order(self):
        return self.target_sh_order

    def get_voxel_size(self):
        """Reference: env.py:478-491."""
        diag = np.diagonal(self.affine_vox2rasmm)[:3]
        return np.mean(np.abs(diag))

    # ------------------------------------------------- stand-alone pieces (tests, diagnostics)
    def _format_state(self, streamlines):
        """Reference: env.py:504-565, for arbitrary streamlines [N,L,3] (numpy or tensor)."""
        pts = torch.as_tensor(np.ascontiguousarray(streamlines, dtype=np.float32)
                              if isinstance(streamlines, np.ndarray) else streamlines)
        pts = pts.to(self.device, dtype=torch.float32).contiguous()
        N, L, _ = pts.shape
        S = self.get_state_size()
        out = torch.zeros((N, S), dtype=torch.float32, device=self.device)
        if N == 0:
            return out
        _lib.check(self._lib.ttl_format_state(ctypes.byref(self._volume), ctypes.byref(self._params),
                                              _lib.ptr(pts), N, L, _lib.ptr(out), S,
                                              _lib.stream_ptr(self.device)), 'ttl_format_state')
        return out

    def _compute_stopping_flags(self, streamlines, with_reward=False):
        """Reference: env.py:567-603.  Returns (should_stop, flags[, mask_value, reward])."""
        pts = torch.as_tensor(np.ascontiguousarray(streamlines, dtype=np.float32)
                              if isinstance(streamlines, np.ndarray) else streamlines)
        pts = pts.to(self.device, dtype=torch.float32).contiguous()
        N, L, _ = pts.shape
        flags = torch.zeros((N,), dtype=torch.int32, device=self.device)
        mval = torch.zeros((N,), dtype=torch.float64, device=self.device)
        rew = torch.zeros((N,), dtype=torch.float32, device=self.device) if with_reward else None
        if N:
            _lib.check(self._lib.ttl_stopping_flags(
                ctypes.byref(self._volume), ctypes.byref(self._params), _lib.ptr(pts), N, L,
                _lib.ptr(flags), _lib.ptr(mval), _lib.ptr(rew), _lib.stream_ptr(self.device)),
                'ttl_stopping_flags')
        f = flags.cpu().numpy().astype(int)
        if with_reward:
            return f != 0, f, mval.cpu().numpy(), rew.cpu().numpy()
        return f != 0, f

    def reset(self):
        pass

    def step(self):
        pass
