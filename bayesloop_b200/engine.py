"""ctypes binding of the C ABI in include/blgrid.h.

The product engine is `bayesloop_b200/csrc/libblgrid.so` (hand-written sm_100a CUDA kernels) driving `cuda`
tensors.  There is NO CPU fallback: if the shared library or a CUDA device is missing, `default_engine()` raises.
PyTorch is used for device memory, streams and torch.distributed only -- every kernel launched on the hot path
comes from libblgrid.so.

`Engine(lib_path, device)` is also the seam the test-suite uses to check the host-side logic of this package on a
machine without a GPU: tests/ build an Engine around the CPU oracle (oracle/libblgrid_oracle.so exports the same
symbols and takes host pointers) and pass it explicitly to the studies.  Nothing in this package refers to oracle/.
"""
import collections
import contextlib
import ctypes
import os
import threading
import weakref

import numpy as np
import torch

MAX_OPS = 16

F_EVIDENCE_ONLY = 1 << 0
F_INIT_STATE = 1 << 1
F_TRANSITION_FIRST = 1 << 2
F_SAVE_STATE = 1 << 3
F_ACCUMULATE = 1 << 4
F_RAW_ALPHA = 1 << 6  # forward rows may stay unnormalised (a backward pass follows)
F_RAW_POSTERIOR = 1 << 7  # backward rows may stay unnormalised; row_scale[b][t] receives the normalising factor
F_NORMALIZE_ROWS = 1 << 5
F_SEPARABLE_ROWS = 1 << 8  # caller's promise about the active operators of every row (include/blgrid.h)

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


class _Problem(ctypes.Structure):
    _fields_ = [('ndim', ctypes.c_int32), ('n', ctypes.c_int32 * 2), ('coords', _dp * 2),
                ('lattice', ctypes.c_double * 2), ('om_kind', ctypes.c_int32), ('seg_len', ctypes.c_int32),
                ('n_cols', ctypes.c_int32), ('reserved', ctypes.c_int32)]


class _Program(ctypes.Structure):
    _fields_ = [('n_ops', ctypes.c_int32), ('kind', ctypes.c_int32 * MAX_OPS), ('axis', ctypes.c_int32 * MAX_OPS),
                ('max_radius', ctypes.c_int32 * MAX_OPS), ('param', ctypes.c_void_p), ('radius', ctypes.c_void_p),
                ('window', ctypes.c_void_p), ('order', ctypes.c_void_p), ('sm_assign', ctypes.c_void_p),
                ('sm_count', ctypes.c_int32), ('sm_slots', ctypes.c_int32)]


class _Inputs(ctypes.Structure):
    _fields_ = [('T', ctypes.c_int64), ('B', ctypes.c_int64), ('data', ctypes.c_void_p), ('prior', ctypes.c_void_p),
                ('reset_base', ctypes.c_void_p), ('lik_table', ctypes.c_void_p), ('prog', _Program),
                ('log_weight', ctypes.c_void_p), ('init_state', ctypes.c_void_p), ('alpha_src', ctypes.c_void_p),
                ('src_stride', ctypes.c_int64)]


class _Outputs(ctypes.Structure):
    _fields_ = [('log_evidence', ctypes.c_void_p), ('local_evidence', ctypes.c_void_p), ('alive', ctypes.c_void_p),
                ('alpha_seq', ctypes.c_void_p), ('avg', ctypes.c_void_p), ('final_state', ctypes.c_void_p),
                ('row_scale', ctypes.c_void_p), ('seq_stride', ctypes.c_int64), ('row_stride', ctypes.c_int64)]


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


_ASSIGNMENTS = {}  # host-side memo of per-SM assignment tables, keyed by the radius table of a sweep


def _give_back(pool, stage):
    if len(pool) < 3:
        pool.append(stage)


class EngineError(RuntimeError):
    pass


class Program:
    """Flat transition program on the engine's device (see transitionModels.LoweringContext)."""

    def __init__(self, engine, ops, B):
        if len(ops) > MAX_OPS:
            raise EngineError('transition program has {} operators, the engine supports {}'.format(len(ops), MAX_OPS))
        self.n_ops = len(ops)
        self.B = B
        self.kinds = [int(o['kind']) for o in ops]
        self.axes = [int(o['axis']) for o in ops]
        K = max(self.n_ops, 1)
        param = np.zeros((B, K))
        radius = np.zeros((B, K), dtype=np.int32)
        window = np.zeros((B, K, 4), dtype=np.int32)
        for k, o in enumerate(ops):
            param[:, k], radius[:, k], window[:, k, :] = o['param'], o['radius'], o['window']
        self.host = dict(param=param, radius=radius, window=window)
        self.param = engine.to_device(param)
        self.radius = engine.to_device(radius)
        self.window = engine.to_device(window)
        self._engine = engine
        self._orders = {}

    def struct(self, lo=0, hi=None):
        hi = self.B if hi is None else hi
        p = _Program()
        p.n_ops = self.n_ops
        for k in range(self.n_ops):
            p.kind[k], p.axis[k] = self.kinds[k], self.axes[k]
            p.max_radius[k] = int(self.host['radius'][lo:hi, k].max()) if hi > lo else 0
        K = max(self.n_ops, 1)
        p.param = self.param.data_ptr() + lo * K * 8
        p.radius = self.radius.data_ptr() + lo * K * 4
        p.window = self.window.data_ptr() + lo * K * 16
        p.order = self._order(lo, hi).data_ptr() if hi - lo > 1 else None
        sms = self._engine.sm_count()
        p.sm_assign, p.sm_count, p.sm_slots = None, 0, 0
        if sms > 0 and hi - lo > sms:
            table = self._assignment(lo, hi, sms)
            if table is not None:
                p.sm_assign, p.sm_count, p.sm_slots = table.data_ptr(), sms, table.shape[1]
        return p

    def _assignment(self, lo, hi, sms, slots=4):
        """Combos of one call pre-assigned to SMs: longest-processing-time-first into `sms` bins of `slots`
        entries.  Cost model fitted to a per-CTA trace of the 1-D kernels on B200 (tools/trace_c2.py, BLG_TRACE;
        1000-cell grid): an SM's time per step grows with the number of resident chains and with the convolution
        work of all of them."""
        if hi - lo > sms * slots:
            return None
        key = (lo, hi, sms, slots)
        if key not in self._orders:
            memo = (self.host['radius'][lo:hi].tobytes(), sms, slots, os.environ.get('BLG_SLOT_ORDER', ''))
            if memo in _ASSIGNMENTS:  # same sweep fitted again (new data, same hyper-grid): reuse the host table
                self._orders[key] = self._engine.to_device(_ASSIGNMENTS[memo])
                return self._orders[key]
            # round 2 (DMMA kernels, fast1d_mma.cuh): the convolution costs ceil((2R + 8 + (R & 1)) / 8) groups of matrix
            # instructions per tile; per-SM trace of a C2 sweep: time ~ 0.17 ms * (chains + groups) per 2000 steps
            rad = self.host['radius'][lo:hi]
            groups = ((2 * rad + 15 + (rad & 1)) // 8).sum(axis=1).astype(float)
            cost = 1.0 + groups
            table = np.full((sms, slots), -1, dtype=np.int32)
            load = np.zeros(sms)
            fill = np.zeros(sms, dtype=np.int64)
            for b in np.argsort(-cost, kind='stable'):
                open_bins = np.flatnonzero(fill < slots)
                sm = open_bins[np.argmin(load[open_bins])]
                table[sm, fill[sm]] = b
                fill[sm] += 1
                load[sm] += cost[b]
            order_mode = os.environ.get('BLG_SLOT_ORDER', 'heavy_first')
            if order_mode != 'heavy_first':
                # CTAs claim the slots of their SM in arrival order and the warp schedulers favour older warps
                for sm in range(sms):
                    k = int(fill[sm])
                    if order_mode == 'light_first':
                        table[sm, :k] = table[sm, :k][::-1].copy()
                    elif order_mode == 'heavy_second' and k >= 2:
                        table[sm, [0, 1]] = table[sm, [1, 0]]
            if len(_ASSIGNMENTS) > 32:
                _ASSIGNMENTS.clear()
            _ASSIGNMENTS[memo] = table
            self._orders[key] = self._engine.to_device(table)
        return self._orders[key]

    def _order(self, lo, hi):
        """Launch order of the combos of one call: most expensive first (cost ~ total convolution radius), so that
        the hardware block scheduler packs heavy and light combinations onto the same SM."""
        if (lo, hi) not in self._orders:
            cost = self.host['radius'][lo:hi].sum(axis=1)
            order = np.argsort(-cost, kind='stable').astype(np.int32)
            self._orders[(lo, hi)] = self._engine.to_device(order)
        return self._orders[(lo, hi)]


class Plan:
    def __init__(self, engine, handle, ndim, n, G):
        self.engine, self.handle, self.ndim, self.n, self.G = engine, handle, ndim, n, G

    def __del__(self):
        try:
            if self.handle:
                self.engine.lib.blg_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Engine:
    """One loaded implementation of include/blgrid.h plus the torch device its pointers live on."""

    def __init__(self, lib_path, device):
        if not os.path.exists(lib_path):
            raise EngineError('engine library not found: {} (run `python -c "import __graft_entry__ as g; g.build()"`)'
                              .format(lib_path))
        self.lib_path = lib_path
        self.device = torch.device(device)
        self.lib = ctypes.CDLL(lib_path)
        L = self.lib
        L.blg_version.restype = ctypes.c_int
        L.blg_last_error.restype = ctypes.c_char_p
        L.blg_backend.restype = ctypes.c_char_p
        L.blg_launch_count.restype = ctypes.c_int64
        L.blg_last_kernel.restype = ctypes.c_char_p
        L.blg_plan_create.argtypes = [ctypes.POINTER(_Problem), ctypes.POINTER(ctypes.c_void_p)]
        L.blg_plan_destroy.argtypes = [ctypes.c_void_p]
        L.blg_plan_destroy.restype = None
        L.blg_plan_set_option.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]
        for name in ('blg_forward', 'blg_backward', 'blg_accumulate'):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.POINTER(_Inputs), ctypes.POINTER(_Outputs),
                                         ctypes.c_uint32, ctypes.c_void_p]
        L.blg_finalize.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                   ctypes.c_uint32, ctypes.c_void_p]
        L.blg_mix.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                              ctypes.c_void_p, ctypes.c_void_p]
        L.blg_wave_weights.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        L.blg_rebase.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p]
        L.blg_share_apply.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Inputs), ctypes.POINTER(_Outputs), ctypes.c_void_p,
                                      ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        L.blg_marginal.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p,
                                   ctypes.c_void_p]
        L.blg_time_average.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        L.blg_scale.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p]
        if L.blg_version() != 1:
            raise EngineError('ABI version mismatch in {}'.format(lib_path))
        self.backend = L.blg_backend().decode()
        self._pinned = {}
        self._plans = collections.OrderedDict()  # problem signature -> Plan (scratch and tables stay allocated)
        self._options = {}  # dispatch options applied to every plan this engine creates (blg_plan_set_option)
        self._free_probe = None  # (torch reservation, driver-reported free bytes) of the last cudaMemGetInfo

    # ---------------------------------------------------------------------------------------------- memory
    def to_device(self, array, pinned=False):
        t = torch.from_numpy(np.ascontiguousarray(array))
        if self.device.type == 'cuda':
            if pinned:
                t = t.pin_memory()
            return t.to(self.device, non_blocking=pinned)
        return t.clone()

    def empty(self, shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def full(self, shape, value, dtype=torch.float64):
        return torch.full(shape if isinstance(shape, (tuple, list)) else (shape,), value, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype=torch.float64):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def to_host(self, tensor):
        """Device -> host.  Large results are downloaded straight into page-locked buffers that the returned NumPy
        array owns; when the array (and every view of it) is garbage collected the buffer goes back to a small pool,
        so steady-state fits neither call cudaHostAlloc (tens of ms for 80 MB) nor fault in fresh pageable memory."""
        t = tensor.detach()
        nbytes = t.numel() * t.element_size()
        if self.device.type == 'cuda' and nbytes >= (1 << 20):
            pool = self._pinned.setdefault(nbytes, [])
            stage = pool.pop() if pool else torch.empty(nbytes, dtype=torch.uint8, device='cpu', pin_memory=True)
            view = stage.view(t.dtype).view(t.shape)
            view.copy_(t.contiguous())
            out = view.numpy()
            weakref.finalize(out, _give_back, pool, stage)
            return out
        return t.cpu().numpy()

    def sm_count(self):
        if self.device.type == 'cuda':
            return torch.cuda.get_device_properties(self.device).multi_processor_count
        return 0

    def free_bytes(self):
        """HBM available to a sweep: what the driver reports free PLUS what torch's caching allocator holds without
        using it (the buffers of the previous fit live there: without them a second fit in the same process would size
        its waves for a nearly full device)."""
        if self.device.type == 'cuda':
            reserved = torch.cuda.memory_reserved(self.device)
            cached = reserved - torch.cuda.memory_allocated(self.device)
            # cudaMemGetInfo costs 2 - 20 ms with tens of GB allocated (tools/e2e_breakdown.py: the largest host-side item
            # of a repeated fit): the driver's figure is only asked again when torch's reservation has changed
            if self._free_probe is None or self._free_probe[0] != reserved:
                self._free_probe = (reserved, torch.cuda.mem_get_info(self.device)[0])
            return self._free_probe[1] + max(0, cached)
        return 8 << 30

    def stream(self):
        if self.device.type == 'cuda':
            return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return ctypes.c_void_p(0)

    def synchronize(self):
        if self.device.type == 'cuda':
            torch.cuda.synchronize(self.device)

    def launch_count(self):
        return int(self.lib.blg_launch_count())

    def last_kernel(self):
        """Kernel family of the last forward / backward pass (tests assert the device path with it)."""
        return self.lib.blg_last_kernel().decode()

    def _check(self, rc):
        if rc != 0:
            raise EngineError('{}: {}'.format(self.backend, self.lib.blg_last_error().decode()))

    # ---------------------------------------------------------------------------------------------- options
    @contextlib.contextmanager
    def options(self, **kw):
        """Dispatch options (include/blgrid.h: blg_plan_set_option) for the plans created inside the `with` block,
        e.g. `with engine.options(force_stream=1): study.fit()`.  They are part of the plan-cache key, so a plan made
        under other options is never reused.  Options choose the kernel family, not the result."""
        saved = dict(self._options)
        self._options.update({k: int(v) for k, v in kw.items()})
        try:
            yield self
        finally:
            self._options = saved

    # ---------------------------------------------------------------------------------------------- calls
    def plan(self, coords, lattice, om_kind, seg_len, n_cols):
        """Plan for one (grid, observation model) description.  Plans are cached per engine (LRU of 8): repeated fits
        of the same problem reuse the device tables and the scratch the library grew inside the plan, so a
        steady-state fit performs no cudaMalloc/cudaFree (each of them synchronises the device)."""
        ndim = len(coords)
        hostCoords = [np.ascontiguousarray(c, dtype=np.float64) for c in coords]
        key = (ndim, tuple(c.tobytes() for c in hostCoords), tuple(float(x) for x in lattice[:ndim]), int(om_kind),
               int(seg_len), int(n_cols), self.stream().value, tuple(sorted(self._options.items())))
        cached = self._plans.get(key)
        if cached is not None:
            self._plans.move_to_end(key)
            return cached
        pb = _Problem()
        pb.ndim = ndim
        for a in range(2):
            pb.n[a] = len(hostCoords[a]) if a < ndim else 1
            pb.lattice[a] = float(lattice[a]) if a < ndim else 1.0
            pb.coords[a] = hostCoords[a].ctypes.data_as(_dp) if a < ndim else None
        pb.om_kind, pb.seg_len, pb.n_cols = int(om_kind), int(seg_len), int(n_cols)
        handle = ctypes.c_void_p()
        if self.device.type == 'cuda':
            with torch.cuda.device(self.device):
                self._check(self.lib.blg_plan_create(ctypes.byref(pb), ctypes.byref(handle)))
        else:
            self._check(self.lib.blg_plan_create(ctypes.byref(pb), ctypes.byref(handle)))
        for name, value in self._options.items():
            self._check(self.lib.blg_plan_set_option(handle, name.encode(), int(value)))
        n = [pb.n[0], pb.n[1]]
        plan = Plan(self, handle, ndim, n, n[0] * n[1])
        self._plans[key] = plan
        while len(self._plans) > 8:
            self._plans.popitem(last=False)
        return plan

    def _io(self, T, B, data, prior, reset_base, lik_table, program, lo, log_weight, init_state, log_evidence,
            local_evidence, alive, alpha_seq, avg, final_state, row_scale=None, seq_stride=0, row_stride=0, alpha_src=None,
            src_stride=0):
        i = _Inputs()
        i.T, i.B = int(T), int(B)
        i.data, i.prior, i.reset_base, i.lik_table = _ptr(data), _ptr(prior), _ptr(reset_base), _ptr(lik_table)
        i.prog = program.struct(lo, lo + B)
        i.log_weight, i.init_state = _ptr(log_weight), _ptr(init_state)
        i.alpha_src, i.src_stride = _ptr(alpha_src), int(src_stride or 0)
        o = _Outputs()
        o.log_evidence, o.local_evidence, o.alive = _ptr(log_evidence), _ptr(local_evidence), _ptr(alive)
        o.alpha_seq, o.avg, o.final_state = _ptr(alpha_seq), _ptr(avg), _ptr(final_state)
        o.row_scale = _ptr(row_scale)
        o.seq_stride, o.row_stride = int(seq_stride or 0), int(row_stride or 0)
        return i, o

    def run(self, which, plan, flags, **kw):
        """which in {'forward', 'backward', 'accumulate'}; keyword arguments are the fields of blg_inputs/outputs
        (tensors on self.device) plus `program` and `lo` (first combo row of the program used by this call)."""
        names = ('T', 'B', 'data', 'prior', 'reset_base', 'lik_table', 'program', 'lo', 'log_weight', 'init_state',
                 'log_evidence', 'local_evidence', 'alive', 'alpha_seq', 'avg', 'final_state', 'row_scale', 'seq_stride',
                 'row_stride', 'alpha_src', 'src_stride')
        args = [kw.get(k) for k in names]
        args[7] = args[7] or 0
        i, o = self._io(*args)
        fn = getattr(self.lib, 'blg_' + which)
        self._check(fn(plan.handle, ctypes.byref(i), ctypes.byref(o), ctypes.c_uint32(flags), self.stream()))

    def finalize(self, plan, seq, T, means, flags):
        self._check(self.lib.blg_finalize(plan.handle, _ptr(seq), int(T), _ptr(means), ctypes.c_uint32(flags),
                                          self.stream()))

    def mix(self, plan, state, weight, K, n, out):
        self._check(self.lib.blg_mix(plan.handle, _ptr(state), _ptr(weight), int(K), int(n), _ptr(out),
                                     self.stream()))

    def wave_weights(self, plan, log_evidence, log_prior, B, shift, avg, count, log_weight):
        """Device-side averaging weights of one wave + re-base of the running sum (include/blgrid.h)."""
        self._check(self.lib.blg_wave_weights(plan.handle, _ptr(log_evidence), _ptr(log_prior), int(B), _ptr(shift),
                                              _ptr(avg), int(count) if avg is not None else 0, _ptr(log_weight),
                                              self.stream()))

    def rebase(self, plan, x, count, shift_from, shift_to):
        self._check(self.lib.blg_rebase(plan.handle, _ptr(x), int(count), _ptr(shift_from), _ptr(shift_to),
                                        self.stream()))

    def share_apply(self, plan, ratio, ratio_stride, n_groups, cp_step, **kw):
        """blg_share_apply: keyword arguments as in run() (T, B, data / lik_table, program, alpha_seq window, strides ...)."""
        names = ('T', 'B', 'data', 'prior', 'reset_base', 'lik_table', 'program', 'lo', 'log_weight', 'init_state',
                 'log_evidence', 'local_evidence', 'alive', 'alpha_seq', 'avg', 'final_state', 'row_scale', 'seq_stride',
                 'row_stride', 'alpha_src', 'src_stride')
        args = [kw.get(k) for k in names]
        args[7] = args[7] or 0
        i, o = self._io(*args)
        self._check(self.lib.blg_share_apply(plan.handle, ctypes.byref(i), ctypes.byref(o), _ptr(ratio), int(ratio_stride),
                                             int(n_groups), _ptr(cp_step), self.stream()))

    def marginal(self, plan, seq, T, axis, out):
        self._check(self.lib.blg_marginal(plan.handle, _ptr(seq), int(T), int(axis), _ptr(out), self.stream()))

    def time_average(self, plan, seq, T, out):
        self._check(self.lib.blg_time_average(plan.handle, _ptr(seq), int(T), _ptr(out), self.stream()))

    def scale(self, plan, x, count, factor):
        self._check(self.lib.blg_scale(plan.handle, _ptr(x), int(count), float(factor), self.stream()))


_lock = threading.Lock()
_default = None


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc', 'libblgrid.so')


def default_engine():
    """The CUDA engine on the current device (LOCAL_RANK aware).  Raises if it cannot be had -- by design."""
    global _default
    with _lock:
        if _default is None:
            if not torch.cuda.is_available():
                raise EngineError('bayesloop_b200 needs a CUDA device (sm_100a); there is no CPU fallback.')
            index = int(os.environ.get('LOCAL_RANK', torch.cuda.current_device()))
            torch.cuda.set_device(index)
            _default = Engine(library_path(), 'cuda:{}'.format(index))
            if not _default.backend.startswith('cuda'):
                raise EngineError('unexpected engine backend {!r}'.format(_default.backend))
        return _default


def set_default_engine(engine):
    """Install `engine` as the process-wide default (used by multi-process launchers and by the test-suite)."""
    global _default
    with _lock:
        _default = engine
