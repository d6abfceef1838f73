"""Device-side engine of the trajectory refinement (host side of csrc/refine.cu).

``RefineEngine`` owns the frame-sharded state of one rank (trajectory with a two-frame halo, Adam moments,
best snapshot, gradient scratch, camera-0 means / inverse covariances, the double control block) and drives
the optimiser steps: ``run`` hands whole iterations to the library (``mc3d_refine_run_*`` picks the two-phase
persistent kernel or the graph of three kernels, see ``plan``), ``one_step`` drives the three phases one by one.
With ``world_size > 1`` (``torch.distributed`` initialised, one process per GPU) frames are sharded contiguously
and the ONLY traffic per step of the three-phase formulation is

    after phase 2 (previous step):  x halo   -- 2 frames to each neighbour (all_gather of the boundary frames)
    after phase 0:                  all-reduce of 7 doubles (cost sums and counts)
    after phase 1:                  all-reduce of 1 double (sum g^2)

(the two-phase step exchanges one block of 13 sums instead of the two all-reduces)

so every rank takes identical clip / Adam / early-stopping decisions with no further communication
(SURVEY.md section 8e).  On GPUs that exchange happens INSIDE the kernels over NVLink peer memory (``PeerExchange``,
csrc/refine.cu ``xchg_*``): every rank maps every other rank's exchange block through CUDA IPC, partial sums and
boundary frames are stored straight into the peers' memory followed by a sequence flag, and consumers spin on flags in
their own memory -- no NCCL call and no host work per step, so any number of steps replays from one CUDA graph.
``torch.distributed`` is only used once, to hand the IPC handles round and for the initial halo.  The host-driven
variant (collectives injected through ``comm``) remains for CPU tensors -- the gloo tests of the sharding logic -- and
as ``MC3D_REFINE_PEER=0``.
"""
import ctypes
import math
import os

import numpy as np

from . import _lib
from . import utils as _utils


def frame_shard(n_frames, rank, world):
    """Contiguous, balanced [begin, end) of ``n_frames`` for ``rank`` of ``world`` (first ranks get the remainder)."""
    base, rem = divmod(n_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def bone_tables(body_lengths, n_joints, connectivity_type='coco'):
    """yaml bones -> (start[], end[], length[]) in yaml key order, plus the joint->bones CSR adjacency.
    Bone names are ``<start>_<end>`` from utils.POINT_INFO (utils.py:1175-1181); an unknown name raises
    KeyError exactly like upstream's ``BPL[bl]`` lookup (pose_refinement.py:851)."""
    bones = _utils.CONNECTIVITY_DICT[connectivity_type]
    names = _utils.generate_connectivity_names(bones, _utils.POINT_INFO[connectivity_type])
    by_name = {names[i]: bones[i] for i in range(len(bones))}
    start, end, length = [], [], []
    for key, val in body_lengths.items():
        s, e = by_name[key]
        if max(s, e) >= n_joints:
            raise IndexError(f'bone {key} needs joint {max(s, e)} but the trajectory has {n_joints} joints')
        start.append(s)
        end.append(e)
        length.append(float(val))
    if len(start) > _lib.MAX_BONES:
        raise ValueError(f'at most {_lib.MAX_BONES} bones are supported')
    adj = [[] for _ in range(n_joints)]
    for k, (s, e) in enumerate(zip(start, end)):
        adj[e].append((k, +1))
        adj[s].append((k, -1))
    adj_start, adj_bone, adj_sign = [0], [], []
    for j in range(n_joints):
        for k, sg in adj[j]:
            adj_bone.append(k)
            adj_sign.append(sg)
        adj_start.append(len(adj_bone))
    return start, end, length, adj_start, adj_bone, adj_sign


def camera_rows(cam_params, camera_ids):
    """[K9 R9 T3 dist5] rows (float64) for the likelihood cameras.  R may be a rotation matrix or an axis-angle
    vector (utils.rotation_conversion, pose_refinement.py:114)."""
    rows = []
    for cid in camera_ids:
        K, R, T, dist = cam_params[cid]
        R = np.asarray(_as_numpy(R), dtype=np.float64)
        if R.shape != (3, 3):
            R = np.asarray(_utils.rotation_conversion(R.reshape(3), to_vector=False), dtype=np.float64)
        d = np.zeros(5)
        dd = np.asarray(_as_numpy(dist), dtype=np.float64).reshape(-1)
        d[:min(5, dd.size)] = dd[:5]
        rows.append(np.concatenate([np.asarray(_as_numpy(K), dtype=np.float64).reshape(9), R.reshape(9),
                                    np.asarray(_as_numpy(T), dtype=np.float64).reshape(3), d]))
    return np.stack(rows)


def _as_numpy(a):
    return a.detach().cpu().numpy() if hasattr(a, 'detach') else np.asarray(a)


class LocalComm:
    """Single-process stand-in for the collectives."""
    rank, world = 0, 1

    def all_gather_object(self, obj):
        return [obj]

    def barrier(self):
        pass

    def all_reduce_sum(self, t):
        return t

    def exchange_halo(self, x_ext, n_local):
        pass

    def all_gather_frames(self, local, total_frames):
        return local


class DistComm:
    """torch.distributed collectives for the frame-sharded refinement (NCCL on GPUs, gloo on CPU tensors)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)

    # The halo traffic is a few hundred bytes per rank, so it rides on all_gather (a plain collective that CUDA
    # graphs capture reliably) instead of point-to-point sends: every rank contributes its two first and two last
    # frames and picks its neighbours' out of the gathered block.
    def exchange_halo(self, x_ext, n_local):
        """x_ext (n_local + 4, J, 3): fill the two halo frames at each end with the neighbours' boundary frames."""
        torch = self._torch()
        mine = torch.cat([x_ext[2:4], x_ext[n_local:n_local + 2]], dim=0).contiguous()        # (4, J, 3)
        flat = torch.empty((self.world * 4,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        self.dist.all_gather_into_tensor(flat, mine, group=self.group)
        out = flat.view((self.world, 4) + tuple(mine.shape[1:]))
        if self.rank > 0:
            x_ext[0:2] = out[self.rank - 1, 2:4]
        if self.rank < self.world - 1:
            x_ext[n_local + 2:n_local + 4] = out[self.rank + 1, 0:2]

    @staticmethod
    def _torch():
        import torch
        return torch

    def all_gather_frames(self, local, total_frames):
        import torch
        sizes = [frame_shard(total_frames, r, self.world) for r in range(self.world)]
        longest = max(e - b for b, e in sizes)
        pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad, group=self.group)
        return torch.cat([o[:e - b] for o, (b, e) in zip(out, sizes)], dim=0)


class _DeviceBytes:
    """Library-owned device memory exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False), 'version': 2}


class PeerExchange:
    """One rank's exchange allocation (mc3d_refine_xchg header + the halo-extended trajectory) and the peers' mappings.

    Layout and protocol: include/mc3d.h (``mc3d_refine_xchg``), csrc/refine.cu (``xchg_publish`` / ``xchg_gather``)."""

    def __init__(self, comm, device, x_bytes):
        import torch
        self.torch, self.comm, self.device = torch, comm, device
        self.lib = _lib.lib()
        self.nbytes = (_lib.XCHG_X_OFFSET + int(x_bytes) + 255) // 256 * 256
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        with torch.cuda.device(device):
            _lib.check(self.lib.mc3d_peer_alloc(self.nbytes, ctypes.byref(ptr), handle))
            self.local_ptr = ptr.value
            self.bytes_view = torch.as_tensor(_DeviceBytes(self.local_ptr, self.nbytes), device=device)
            self.ptrs = [None] * comm.world
            self.ptrs[comm.rank] = self.local_ptr
            self._opened = []
            self.failed = None
            if comm.world > 1:
                if comm.world > _lib.MAX_PEERS:
                    raise ValueError(f'at most {_lib.MAX_PEERS} ranks are supported by the in-kernel exchange')
                handles = comm.all_gather_object(bytes(handle))
                try:
                    if os.environ.get('MC3D_PEER_FAIL') == '1':            # test hook: behave as if a peer could not be mapped
                        raise _lib.Mc3dError('MC3D_PEER_FAIL=1')
                    for r, h in enumerate(handles):
                        if r == comm.rank:
                            continue
                        p = ctypes.c_void_p()
                        hb = (ctypes.c_ubyte * _lib.IPC_HANDLE_BYTES).from_buffer_copy(h)
                        _lib.check(self.lib.mc3d_peer_open(hb, ctypes.byref(p)))
                        self.ptrs[r] = p.value
                        self._opened.append(p.value)
                except _lib.Mc3dError as exc:                               # e.g. the peer's GPU is not visible to this process
                    self.failed = str(exc)
                # every rank must take the same path: one failure sends all of them to the host-driven exchange
                reasons = [r for r in comm.all_gather_object(self.failed) if r]
                if reasons:
                    self.failed = reasons[0]
                    self.close()

    def x_view(self, shape, dtype):
        """The trajectory tensor living at byte XCHG_X_OFFSET of the local allocation."""
        n = int(np.prod(shape)) * self.torch.empty((), dtype=dtype).element_size()
        return self.bytes_view[_lib.XCHG_X_OFFSET:_lib.XCHG_X_OFFSET + n].view(dtype).view(*shape)

    def error(self):
        """True when a wait on a peer timed out in some kernel of this rank."""
        head = self.bytes_view[:ctypes.sizeof(_lib.RefineXchg)].cpu().numpy().tobytes()
        return _lib.RefineXchg.from_buffer_copy(head).error != 0

    def close(self):
        """Collective: nobody unmaps or frees while a peer may still store into the block."""
        if self.local_ptr is None:
            return
        with self.torch.cuda.device(self.device):
            self.torch.cuda.synchronize()
            self.comm.barrier()
            for p in self._opened:
                self.lib.mc3d_peer_close(p)
            self._opened = []
            self.comm.barrier()
            self.bytes_view = None
            self.lib.mc3d_peer_free(self.local_ptr)
            self.local_ptr = None


class CudaPhases:
    """The three phases on the GPU through the C ABI (mc3d_refine_phase_* / mc3d_refine_run_*)."""

    def __init__(self, dtype_tag):
        lib = _lib.lib()
        self.phase_fn = getattr(lib, f'mc3d_refine_phase_{dtype_tag}')
        self.run_fn = getattr(lib, f'mc3d_refine_run_{dtype_tag}')
        self.flags_fn = getattr(lib, f'mc3d_refine_flags_{dtype_tag}')

    def phase(self, problem, phase, step, end_of_iteration, stream):
        _lib.check(self.phase_fn(ctypes.byref(problem), phase, step, int(end_of_iteration), stream))

    def flags(self, problem, stream):
        _lib.check(self.flags_fn(ctypes.byref(problem), stream))

    def run(self, problem, first_step, n_iters, stream):
        _lib.check(self.run_fn(ctypes.byref(problem), first_step, n_iters, stream))


class RefineEngine:
    """Frame-sharded optimiser state of one rank plus the step driver."""

    def __init__(self, trajectory, gaussians, cam_rows, body_lengths, *, torch_dtype, device, lr, betas,
                 lambda_smooth, lambda_body_length, patience, tolerance, max_iter, ignore_distortions,
                 window, n_window_frames, hist_capacity, comm=None, phases=None, gaussian_camera=0, adam_eps=1e-8,
                 use_exchange=None, gaussian_cameras=None):
        """``gaussian_cameras``: None = camera ``gaussian_camera``'s Gaussians for every camera (upstream's behaviour, quirk Q1);
        a list with one entry per row of ``cam_rows`` = per-camera Gaussians (camera c is compared with
        ``gaussians[:, gaussian_cameras[c]]``)."""
        import torch
        self.torch = torch
        self.comm = comm or LocalComm()
        self.dtype = torch_dtype
        self.tag = 'f32' if torch_dtype == torch.float32 else 'f64'
        if torch_dtype not in (torch.float32, torch.float64):
            raise TypeError('torch_dtype must be float32 or float64')
        self.device = torch.device(device)
        if self.device.type != 'cuda' and phases is None:
            # CPU tensors are accepted only together with injected phases (the gloo tests of the sharding logic)
            raise _lib.Mc3dError('RefineEngine needs a CUDA device: the refinement kernels have no CPU fallback')
        self.total_frames = int(trajectory.shape[0])
        self.J = int(trajectory.shape[1])
        if self.J > _lib.MAX_JOINTS:
            raise ValueError(f'at most {_lib.MAX_JOINTS} joints are supported')
        self.begin, self.end = frame_shard(self.total_frames, self.comm.rank, self.comm.world)
        n = self.n_local = self.end - self.begin
        dev, dt = self.device, torch_dtype

        traj = torch.as_tensor(trajectory).to(dt)
        # GPUs: the kernels exchange sums and halo frames themselves over peer memory, and the trajectory lives in
        # the exchange allocation (one GPU: the block only serves the persistent kernel's grid barriers).
        # MC3D_REFINE_PEER=0 selects the host-driven exchange / the plain three-kernel graph instead.
        mode = os.environ.get('MC3D_REFINE_PEER', '')
        self.peer = None
        if dev.type == 'cuda' and phases is None and mode != '0' and use_exchange is not False:
            if self.comm.world > 1 and min(frame_shard(self.total_frames, r, self.comm.world)[1] -
                                           frame_shard(self.total_frames, r, self.comm.world)[0]
                                           for r in range(self.comm.world)) < 2:
                raise ValueError('the frame-sharded refinement needs at least two frames per rank')
            self.peer = PeerExchange(self.comm, dev, (n + 4) * self.J * 3 * torch.empty((), dtype=dt).element_size())
            if self.peer.failed:
                import warnings
                warnings.warn(f'in-kernel exchange unavailable ({self.peer.failed}); using the host-driven exchange')
                self.peer = None
        if self.peer is not None:
            self.x_ext = self.peer.x_view((n + 4, self.J, 3), dt)
        else:
            self.x_ext = torch.zeros((n + 4, self.J, 3), dtype=dt, device=dev)
        self.x_ext[2:n + 2] = traj[self.begin:self.end].to(dev)
        self.m = torch.zeros((n, self.J, 3), dtype=dt, device=dev)
        self.v = torch.zeros_like(self.m)
        self.best = torch.zeros_like(self.m)
        self.g = torch.zeros_like(self.m)
        # two-phase step (csrc/refine.cu costgrad_loop): likelihood, smoothness and the two bone-length components
        # (four components at a stride of n*J*3 rounded up to a multiple of 4 scalars)
        self.gc = torch.zeros((4 * ((n * self.J * 3 + 3) // 4 * 4),), dtype=dt, device=dev) if self.peer is not None else None
        self.term_ok = torch.zeros((n + 4,), dtype=torch.uint8, device=dev)
        self.gaussian_cameras = None if gaussian_cameras is None else [int(c) for c in gaussian_cameras]
        if self.gaussian_cameras is not None and len(self.gaussian_cameras) != int(cam_rows.shape[0]):
            raise ValueError('gaussian_cameras needs one entry per camera')
        n_gc = 1 if self.gaussian_cameras is None else len(self.gaussian_cameras)
        self.mu0 = torch.zeros((n_gc, n, self.J, 2), dtype=dt, device=dev)     # (cameras, frames, joints, .): one camera unless per-camera
        self.S = torch.zeros((n_gc, n, self.J, 3), dtype=dt, device=dev)
        self.hist_capacity = int(hist_capacity)
        self.ctrl = torch.zeros((_lib.CT_HIST + 4 * self.hist_capacity,), dtype=torch.float64, device=dev)
        self.ctrl[_lib.CT_STATE + 3] = math.inf
        self.ctrl[_lib.CT_STATE + 16 + 3] = math.inf

        # camera-0 means and inverse covariances for my frames (frames beyond the gaussians stay zero: never in a window)
        g_all = torch.as_tensor(gaussians).to(dt)
        n_g = max(0, min(self.end, int(g_all.shape[0])) - self.begin)
        self.n_cams_in_gaussians = int(g_all.shape[1])
        if n_g > 0:
            g_loc = g_all[self.begin:self.begin + n_g].to(dev).contiguous()
            for slot, gcam in enumerate([gaussian_camera] if self.gaussian_cameras is None else self.gaussian_cameras):
                if dev.type == 'cuda':
                    fn = getattr(_lib.lib(), f'mc3d_refine_prepare_{self.tag}')
                    with torch.cuda.device(dev):
                        _lib.check(fn(g_loc.data_ptr(), n_g, int(g_loc.shape[1]), self.J, int(gcam), 1e-6,
                                      self.mu0[slot].data_ptr(), self.S[slot].data_ptr(), torch.cuda.current_stream().cuda_stream))
                        torch.cuda.current_stream().synchronize()           # g_loc may be freed after this
                else:
                    self._prepare_host(g_loc, n_g, gcam, slot)

        start, end_, length, adj_start, adj_bone, adj_sign = bone_tables(body_lengths, self.J)
        pb = self.problem = _lib.RefineProblem()
        pb.n_joints, pb.n_cams, pb.n_bones = self.J, int(cam_rows.shape[0]), len(start)
        if pb.n_cams > _lib.MAX_VIEWS:
            raise ValueError(f'at most {_lib.MAX_VIEWS} cameras are supported')
        pb.ignore_distortions = int(bool(ignore_distortions))
        pb.patience = int(min(patience, 2 ** 31 - 1)) if math.isfinite(patience) else 2 ** 31 - 1
        pb.max_iter = int(min(max_iter, 2 ** 31 - 2)) if math.isfinite(max_iter) else 2 ** 31 - 2
        pb.n_frames, pb.frame_offset = n, self.begin
        pb.win_begin, pb.win_end = int(window[0]), int(window[1])
        pb.hist_capacity = self.hist_capacity
        pb.total_frames = self.total_frames
        pb.lr, pb.beta1, pb.beta2, pb.eps = float(lr), float(betas[0]), float(betas[1]), float(adam_eps)
        pb.lambda_smooth, pb.lambda_body = float(lambda_smooth), float(lambda_body_length)
        pb.tolerance = float(tolerance)
        pb.aa = float(n_window_frames) * float(sum(a * a for a in length))
        for c in range(pb.n_cams):
            for i in range(26):
                pb.cams[c][i] = float(cam_rows[c, i])
        for k in range(len(start)):
            pb.bone_start[k], pb.bone_end[k], pb.bone_len[k] = start[k], end_[k], length[k]
        for i, a in enumerate(adj_start):
            pb.adj_start[i] = a
        for i, (b, sg) in enumerate(zip(adj_bone, adj_sign)):
            pb.adj_bone[i], pb.adj_sign[i] = b, sg
        for name in ('m', 'v', 'best', 'g', 'mu0', 'S', 'term_ok', 'ctrl'):
            setattr(pb, name, getattr(self, name).data_ptr())
        pb.x = self.x_ext.data_ptr()
        pb.gauss_cam_stride = 0 if self.gaussian_cameras is None else n * self.J
        pb.test_flags = int(os.environ.get('MC3D_REFINE_TEST_FLAGS', '0'))      # test hook (include/mc3d.h)
        if self.peer is not None:
            pb.gc = self.gc.data_ptr()
            pb.rank, pb.world = self.comm.rank, self.comm.world
            if self.comm.rank > 0:
                lb, le = frame_shard(self.total_frames, self.comm.rank - 1, self.comm.world)
                pb.n_frames_left = le - lb
            pb.spin_timeout_ns = int(float(os.environ.get('MC3D_REFINE_SPIN_TIMEOUT_S', '10')) * 1e9)
            for r, p in enumerate(self.peer.ptrs):
                pb.xchg[r] = p
        self.phases = phases or CudaPhases(self.tag)
        self.step = 0
        self.use_graph = os.environ.get('MC3D_REFINE_GRAPH', '1') != '0'      # multi-rank CUDA-graph replay
        self._graph = None
        if self.comm.world > 1:
            self.comm.exchange_halo(self.x_ext, n)
        if getattr(self.phases, 'engine', 0) is None:         # test phases operate on this engine's tensors
            self.phases.engine = self
        self.phases.flags(self.problem, self._stream())       # smoothness-term validity, once per run

    def _prepare_host(self, g_loc, n_g, cam, slot=0):
        # CPU tensors exist only for the gloo tests of the sharding logic (tests inject their own phases).
        torch = self.torch
        gp = g_loc[:, cam]
        c00, c01, c10, c11 = gp[..., 2] + 1e-6, gp[..., 3], gp[..., 4], gp[..., 5] + 1e-6
        det = c00.double() * c11.double() - c01.double() * c10.double()
        self.mu0[slot, :n_g] = gp[..., :2]
        self.S[slot, :n_g] = torch.stack([c11.double() / det, -0.5 * (c01.double() + c10.double()) / det,
                                    c00.double() / det], dim=-1).to(self.dtype)

    # ---- stepping ------------------------------------------------------------------------------------------------
    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream if self.device.type == 'cuda' else None

    def set_window(self, begin, end):
        self.problem.win_begin, self.problem.win_end = int(begin), int(end)

    def one_step(self, end_of_iteration=True):
        """Phases 0, 1, 2 with the inter-rank exchanges in between."""
        p = self.step & 1
        acc = self.ctrl[_lib.CT_ACC + 16 * p:_lib.CT_ACC + 16 * p + 8]
        st = self._stream()
        host_exchange = self.comm.world > 1 and self.peer is None
        self.phases.phase(self.problem, 0, self.step, end_of_iteration, st)
        if host_exchange:
            self.comm.all_reduce_sum(acc[0:7])
        self.phases.phase(self.problem, 1, self.step, end_of_iteration, st)
        if host_exchange:
            self.comm.all_reduce_sum(acc[7:8])
        self.phases.phase(self.problem, 2, self.step, end_of_iteration, st)
        if host_exchange:
            self.comm.exchange_halo(self.x_ext, self.n_local)
        self.step += 1

    def run(self, n_iters):
        """``n_iters`` whole-window iterations.

        One rank, or several ranks with the in-kernel exchange: a single C call replaying a CUDA graph of the three
        kernels (the ranks meet inside the kernels).  Host-driven exchange over NCCL: two consecutive steps --
        kernels, the two all-reduces and the halo exchanges -- are captured once into a CUDA graph and replayed."""
        n_iters = int(n_iters)
        if (self.comm.world == 1 or self.peer is not None) and hasattr(self.phases, 'run') and self.device.type == 'cuda':
            with self.torch.cuda.device(self.device):
                self.phases.run(self.problem, self.step, n_iters, self._stream())
            self.step += n_iters
            return
        done = 0
        if self.comm.world > 1 and self.device.type == 'cuda' and self.use_graph and n_iters >= 8:
            torch = self.torch
            with torch.cuda.device(self.device):
                while done < 2 or (self.step & 1):              # eager warm-up (NCCL channels), even parity for the graph
                    self.one_step(True)
                    done += 1
                if self._graph is None:
                    first = self.step
                    ok = 1.0
                    graph = None
                    try:
                        torch.cuda.synchronize()
                        graph = torch.cuda.CUDAGraph()
                        # thread-local: the NCCL watchdog thread must not invalidate the capture
                        with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                            self.one_step(True)
                            self.one_step(True)
                    except Exception:                            # capture unsupported here
                        ok = 0.0
                    self.step = first                           # capturing did not execute anything
                    torch.cuda.synchronize()
                    flag = torch.tensor([ok], dtype=torch.float64, device=self.device)
                    self.comm.dist.all_reduce(flag, op=self.comm.dist.ReduceOp.MIN, group=self.comm.group)
                    if flag.item() == 1.0:                      # every rank captured: replay; otherwise all stay eager
                        self._graph = graph
                    else:
                        self.use_graph = False
                if self._graph is not None:
                    for _ in range((n_iters - done) // 2):
                        self._graph.replay()
                        self.step += 2
                        done += 2
        for _ in range(n_iters - done):
            self.one_step(True)

    def plan(self):
        """What ``run`` launches for this engine (text from the library)."""
        if self.device.type != 'cuda' or not hasattr(self.phases, 'run_fn'):
            return 'host-driven phases'
        if self.comm.world > 1 and self.peer is None:
            return 'three-phase step, host-driven exchange over torch.distributed (NCCL), 2-step CUDA graph'
        with self.torch.cuda.device(self.device):
            return _lib.lib().mc3d_refine_plan(ctypes.byref(self.problem)).decode()

    def release_graph(self):
        """Drop the captured multi-rank graph (it holds NCCL kernels: release it before the process group goes away)."""
        if self._graph is not None:
            self.torch.cuda.synchronize()
            self._graph = None

    def close(self):
        """Collective when sharded: release the graph and the peer mappings.  The trajectory tensors read back before
        this call stay valid only if they were copied (``trajectory()`` / ``best_trajectory()`` return copies)."""
        self.release_graph()
        if self.peer is not None:
            if self.peer.error():
                self.peer.close()
                self.peer = None
                raise _lib.Mc3dError('a wait on a peer rank timed out inside a refinement kernel (results are invalid)')
            self.x_ext = self.x_ext.clone()
            self.problem.x = self.x_ext.data_ptr()
            for r in range(_lib.MAX_PEERS):
                self.problem.xchg[r] = None
            self.peer.close()
            self.peer = None

    # ---- read-back ---------------------------------------------------------------------------------------------------
    def state(self):
        """State entering the next step: dict(adam_step, best, no_improve, stopped, iterations, improved)."""
        s = self.ctrl[_lib.CT_STATE + 16 * (self.step & 1):_lib.CT_STATE + 16 * (self.step & 1) + 8].cpu().numpy()
        return dict(adam_step=int(s[0]), run_sum=float(s[1]), run_cnt=int(s[2]), best=float(s[3]), no_improve=int(s[4]),
                    stopped=bool(s[5]), iterations=int(s[6]), improved=bool(s[7]))

    def history(self, n_steps):
        """(n_steps, 4) float64 [total, likelihood, smoothness, body_length] per optimiser step."""
        n_steps = min(int(n_steps), self.hist_capacity)
        return self.ctrl[_lib.CT_HIST:_lib.CT_HIST + 4 * n_steps].cpu().numpy().reshape(n_steps, 4)

    def trajectory(self):
        return self.comm.all_gather_frames(self.x_ext[2:self.n_local + 2].clone(), self.total_frames)

    def best_trajectory(self):
        return self.comm.all_gather_frames(self.best, self.total_frames)
