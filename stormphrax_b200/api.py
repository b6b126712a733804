"""ctypes binding of libsp_nnue.so (include/sp_nnue.h) -- the product's Python face.

Everything here calls the C-ABI; there is no Python or CPU implementation of the evaluation
behind it.  If the shared library is missing, or no CUDA device is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

BOARD_DTYPE = np.dtype(
    [
        ("occupancy", "<u8"),
        ("pieces", "u1", (16,)),
        ("stm_ep", "u1"),
        ("halfmove", "u1"),
        ("fullmove", "<u2"),
        ("eval", "<i2"),
        ("wdl", "u1"),
        ("extra", "u1"),
    ]
)
assert BOARD_DTYPE.itemsize == 32

SP_OK, SP_ERR_INVALID, SP_ERR_BAD_NETWORK, SP_ERR_CUDA, SP_ERR_NO_DEVICE, SP_ERR_BAD_BOARD, SP_ERR_CAPACITY = range(7)
STATUS_NAMES = ["SP_OK", "SP_ERR_INVALID", "SP_ERR_BAD_NETWORK", "SP_ERR_CUDA", "SP_ERR_NO_DEVICE", "SP_ERR_BAD_BOARD", "SP_ERR_CAPACITY"]
NUM_COUNTERS = 8
CTR_EVALS, CTR_FULL_REFRESH, CTR_INCREMENTAL, CTR_LAUNCHES = 0, 1, 2, 3
KERNEL_CLASSES = ["ft_full", "head", "ft_slots", "ft_games", "extract", "accumulate", "rebuilds", "head_main"]
NUM_KERNEL_CLASSES = len(KERNEL_CLASSES)

_vp = C.c_void_p
_sz = C.c_size_t

# name -> (restype, argtypes): every symbol include/sp_nnue.h declares
SIGNATURES = {
    "sp_nnue_create": (C.c_int, [_vp, _sz, C.c_int, C.POINTER(_vp)]),
    "sp_nnue_destroy": (None, [_vp]),
    "sp_nnue_last_error": (C.c_char_p, [_vp]),
    "sp_nnue_device": (C.c_int, [_vp]),
    "sp_nnue_sync": (C.c_int, [_vp, _vp]),
    "sp_nnue_set_stream": (C.c_int, [_vp, _vp]),
    "sp_nnue_profile": (C.c_int, [_vp, C.c_int]),
    "sp_nnue_profile_read": (C.c_int, [_vp, _vp, _vp]),
    "sp_host_feature_counts": (C.c_int, [_vp, _sz, C.c_int, _vp]),
    "sp_host_playout_stats": (C.c_int, [_vp, _vp, C.c_uint32, C.c_int, _vp]),
    "sp_nnue_eval_full": (C.c_int, [_vp, _vp, _sz, _vp]),
    "sp_nnue_eval_full_device": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "sp_nnue_slots_reserve": (C.c_int, [_vp, _sz]),
    "sp_nnue_refresh": (C.c_int, [_vp, _vp, _vp, _sz]),
    "sp_nnue_update": (C.c_int, [_vp, _vp, _vp, _vp, _sz]),
    "sp_nnue_eval_slots": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "sp_nnue_update_eval": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "sp_nnue_batch": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "sp_nnue_refresh_device": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "sp_nnue_update_eval_device": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "sp_nnue_eval_playouts": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp]),
    "sp_nnue_eval_playouts_device": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _sz, _vp, _vp]),
    "sp_nnue_forward_device": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "sp_nnue_activations_device": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "sp_nnue_adjust_defaults": (None, [_vp]),
    "sp_nnue_adjust": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "sp_nnue_adjust_device": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "sp_nnue_counters": (C.c_int, [_vp, _vp]),
    "sp_nnue_read_slot": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "sp_host_playouts": (_sz, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp]),
    "sp_host_board_from_fen": (C.c_int, [C.c_char_p, _vp]),
    "sp_host_board_to_fen": (C.c_int, [_vp, C.c_char_p, _sz]),
    "sp_host_board_from_dfrc": (C.c_int, [C.c_uint32, _vp]),
    "sp_host_legal_moves": (C.c_int, [_vp, _vp]),
    "sp_host_in_check": (C.c_int, [_vp]),
    "sp_host_adjust": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "sp_host_apply_move": (C.c_int, [_vp, C.c_uint16, _vp]),
    "sp_host_features": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "sp_host_feature_delta": (C.c_int, [_vp, _vp, C.c_int] + [_vp] * 8),
    "sp_selfplay_run": (C.c_int, [_vp, _sz, C.c_int, _vp, _vp, _vp, _sz, _vp]),
    "sp_selfplay_run_gpu": (C.c_int, [_vp, _sz, C.c_int, _vp, _vp, _vp, _sz, _vp]),
    "sp_nnue_batch_device": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp]),
    "sp_host_viriformat": (C.c_long, [_vp, _vp, _vp, C.c_uint32, C.c_int, _vp, _sz]),
    "sp_host_normalize_score": (C.c_int, [_vp, C.c_int32, _vp, _vp]),
    "sp_host_wdl_model": (C.c_int, [_vp, C.c_int32, _vp, _vp]),
    "sp_host_net_payload": (C.c_long, [_vp, _sz, _vp, _sz]),
    "sp_nnue_wdl": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "sp_nnue_wdl_device": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
}



class AdjustParams(C.Structure):
    """SpAdjustParams (include/sp_nnue.h): tunables of eval::adjustEval plus contempt / optimism."""

    _fields_ = [
        ("scaling_value", C.c_int32 * 5),
        ("material_scaling_base", C.c_int32),
        ("optimism_base", C.c_int32),
        ("optimism_material_scale", C.c_int32),
        ("contempt", C.c_int32 * 2),
        ("optimism", C.c_int32 * 2),
    ]

    @classmethod
    def defaults(cls, contempt=(0, 0), optimism=(0, 0)) -> "AdjustParams":
        p = cls()
        lib().sp_nnue_adjust_defaults(C.byref(p))
        p.contempt[:] = contempt
        p.optimism[:] = optimism
        return p


_lib = None


class NnueError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else status}: {message}")
        self.status = status


def lib() -> C.CDLL:
    """Load (building first if the sources are newer) libsp_nnue.so and bind every export."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        if os.environ.get("SP_NNUE_LIB"):  # experiments: a variant build of the same sources
            path = os.environ["SP_NNUE_LIB"]
        elif os.path.exists(os.path.join(_build.CSRC, "kernels.cu")) and _build.shutil.which("nvcc"):
            path = _build.build()
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing and cannot be built here: run `python -m stormphrax_b200.build`")
        handle = C.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError = the library does not export what the header declares
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def _ptr(a) -> int | None:
    """Host numpy array, torch tensor (host or device) or raw int -> address."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)


def _boards(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=BOARD_DTYPE)


class Nnue:
    """One evaluator context = one GPU holding one copy of the network (eval::init .. shutdown)."""

    def __init__(self, net_image, device: int = 0):
        self._lib = lib()
        image = np.ascontiguousarray(np.frombuffer(net_image, dtype=np.uint8) if isinstance(net_image, (bytes, bytearray)) else net_image, dtype=np.uint8)
        handle = _vp()
        rc = self._lib.sp_nnue_create(image.ctypes.data, image.size, device, C.byref(handle))
        if rc:
            raise NnueError(rc, self._lib.sp_nnue_last_error(None).decode())
        self._h = handle

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.sp_nnue_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc:
            raise NnueError(rc, self._lib.sp_nnue_last_error(self._h).decode())

    @property
    def device(self) -> int:
        return self._lib.sp_nnue_device(self._h)

    def sync(self, stream: int | None = None) -> None:
        self._check(self._lib.sp_nnue_sync(self._h, stream))

    def set_stream(self, stream: int | None) -> None:
        self._check(self._lib.sp_nnue_set_stream(self._h, stream))

    def profile(self, enable: bool) -> None:
        self._check(self._lib.sp_nnue_profile(self._h, int(enable)))

    def profile_read(self) -> dict:
        """{kernel class: (milliseconds, launches)} since the previous read."""
        ms = np.zeros(NUM_KERNEL_CLASSES, dtype=np.float64)
        launches = np.zeros(NUM_KERNEL_CLASSES, dtype=np.uint64)
        self._check(self._lib.sp_nnue_profile_read(self._h, ms.ctypes.data, launches.ctypes.data))
        return {name: (float(ms[i]), int(launches[i])) for i, name in enumerate(KERNEL_CLASSES)}

    def counters(self) -> np.ndarray:
        out = np.zeros(NUM_COUNTERS, dtype=np.uint64)
        self._check(self._lib.sp_nnue_counters(self._h, out.ctypes.data))
        return out

    # ---- full refresh (NnueState::evaluateOnce for a batch)
    def eval_full(self, boards, out: np.ndarray | None = None) -> np.ndarray:
        boards = _boards(boards)
        if out is None:
            out = np.empty(boards.size, dtype=np.int32)
        self._check(self._lib.sp_nnue_eval_full(self._h, boards.ctypes.data, boards.size, out.ctypes.data))
        return out

    def eval_full_device(self, d_boards, n: int, d_out, stream: int | None = None) -> None:
        self._check(self._lib.sp_nnue_eval_full_device(self._h, _ptr(d_boards), n, _ptr(d_out), stream))

    def activations_device(self, d_boards, n: int, d_act, d_bucket, stream: int | None = None) -> None:
        self._check(self._lib.sp_nnue_activations_device(self._h, _ptr(d_boards), n, _ptr(d_act), _ptr(d_bucket), stream))

    def forward_device(self, d_act, d_bucket, n: int, d_out, stream: int | None = None) -> None:
        self._check(self._lib.sp_nnue_forward_device(self._h, _ptr(d_act), _ptr(d_bucket), n, _ptr(d_out), stream))

    # ---- eval post-processing (adjustStatic + adjustEval)
    def adjust(self, boards, raw, params: "AdjustParams | None" = None, correction=None) -> np.ndarray:
        boards = _boards(boards)
        raw = np.ascontiguousarray(raw, dtype=np.int32)
        corr = None if correction is None else np.ascontiguousarray(correction, dtype=np.int32)
        params = params or AdjustParams.defaults()
        out = np.empty(boards.size, dtype=np.int32)
        self._check(self._lib.sp_nnue_adjust(self._h, boards.ctypes.data, raw.ctypes.data, _ptr(corr), boards.size, C.addressof(params), out.ctypes.data))
        return out

    def wdl(self, boards, scores, model: bool = True):
        """(normalizeScore<false>, win per mille, loss per mille) of `scores` on `boards`, computed on the device."""
        boards = _boards(boards)
        scores = np.ascontiguousarray(scores, dtype=np.int32)
        norm = np.empty(boards.size, dtype=np.int32)
        win = np.empty(boards.size, dtype=np.int32) if model else None
        loss = np.empty(boards.size, dtype=np.int32) if model else None
        self._check(self._lib.sp_nnue_wdl(self._h, boards.ctypes.data, scores.ctypes.data, boards.size, norm.ctypes.data, _ptr(win), _ptr(loss)))
        return norm, win, loss

    # ---- accumulator slots (the NnueState stack, device resident)
    def slots_reserve(self, n_slots: int) -> None:
        self._check(self._lib.sp_nnue_slots_reserve(self._h, n_slots))

    def refresh(self, slots, boards) -> None:
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        boards = _boards(boards)
        assert slots.size == boards.size
        self._check(self._lib.sp_nnue_refresh(self._h, slots.ctypes.data, boards.ctypes.data, slots.size))

    def update(self, src_slots, dst_slots, after) -> None:
        src = np.ascontiguousarray(src_slots, dtype=np.uint32)
        dst = np.ascontiguousarray(dst_slots, dtype=np.uint32)
        after = _boards(after)
        assert src.size == dst.size == after.size
        self._check(self._lib.sp_nnue_update(self._h, src.ctypes.data, dst.ctypes.data, after.ctypes.data, src.size))

    def eval_slots(self, slots, stm=None) -> np.ndarray:
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        stm_arr = None if stm is None else np.ascontiguousarray(stm, dtype=np.uint8)
        out = np.empty(slots.size, dtype=np.int32)
        self._check(self._lib.sp_nnue_eval_slots(self._h, slots.ctypes.data, _ptr(stm_arr), slots.size, out.ctypes.data))
        return out

    def update_eval(self, src_slots, dst_slots, after) -> np.ndarray:
        src = np.ascontiguousarray(src_slots, dtype=np.uint32)
        dst = np.ascontiguousarray(dst_slots, dtype=np.uint32)
        after = _boards(after)
        assert src.size == dst.size == after.size
        out = np.empty(src.size, dtype=np.int32)
        self._check(self._lib.sp_nnue_update_eval(self._h, src.ctypes.data, dst.ctypes.data, after.ctypes.data, src.size, out.ctypes.data))
        return out

    def batch(self, refresh=None, update=None, evaluate=None):
        """One round of a batched driver (sp_nnue_batch): refresh = (slots, boards), update = (src, dst, boards),
        evaluate = (slots, stm or None).  Returns (refresh evals, update evals, evaluate evals); absent groups give None."""
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
        r_slots = r_boards = r_out = u_src = u_dst = u_boards = u_out = e_slots = e_stm = e_out = None
        n_r = n_u = n_e = 0
        if refresh is not None:
            r_slots, r_boards = u32(refresh[0]), _boards(refresh[1])
            n_r, r_out = r_slots.size, np.empty(r_slots.size, dtype=np.int32)
        if update is not None:
            u_src, u_dst, u_boards = u32(update[0]), u32(update[1]), _boards(update[2])
            n_u, u_out = u_src.size, np.empty(u_src.size, dtype=np.int32)
        if evaluate is not None:
            e_slots = u32(evaluate[0])
            e_stm = None if evaluate[1] is None else np.ascontiguousarray(evaluate[1], dtype=np.uint8)
            n_e, e_out = e_slots.size, np.empty(e_slots.size, dtype=np.int32)
        self._check(self._lib.sp_nnue_batch(self._h, _ptr(r_slots), _ptr(r_boards), n_r, _ptr(r_out), _ptr(u_src), _ptr(u_dst), _ptr(u_boards), n_u, _ptr(u_out),
                                            _ptr(e_slots), _ptr(e_stm), n_e, _ptr(e_out)))
        return r_out, u_out, e_out

    def refresh_device(self, d_slots, d_boards, n: int, stream: int | None = None) -> None:
        self._check(self._lib.sp_nnue_refresh_device(self._h, _ptr(d_slots), _ptr(d_boards), n, stream))

    def update_eval_device(self, d_src, d_dst, d_after, n: int, d_out, stream: int | None = None) -> None:
        self._check(self._lib.sp_nnue_update_eval_device(self._h, _ptr(d_src), _ptr(d_dst), _ptr(d_after), n, _ptr(d_out), stream))

    def read_slot(self, slot: int):
        acc = np.empty((2, 1024), dtype=np.int16)
        board = np.zeros(1, dtype=BOARD_DTYPE)
        self._check(self._lib.sp_nnue_read_slot(self._h, slot, acc.ctypes.data, board.ctypes.data))
        return acc, board

    # ---- playout streams (datagen form: one accumulator chain per game)
    def eval_playouts(self, boards, game_start) -> np.ndarray:
        boards = _boards(boards)
        game_start = np.ascontiguousarray(game_start, dtype=np.uint32)
        assert game_start[-1] == boards.size
        out = np.empty(boards.size, dtype=np.int32)
        self._check(self._lib.sp_nnue_eval_playouts(self._h, boards.ctypes.data, game_start.ctypes.data, game_start.size - 1, out.ctypes.data))
        return out

    def eval_playouts_device(self, d_boards, d_game_start, n_games: int, n_boards: int, d_out, stream: int | None = None) -> None:
        self._check(self._lib.sp_nnue_eval_playouts_device(self._h, _ptr(d_boards), _ptr(d_game_start), n_games, n_boards, _ptr(d_out), stream))


# ------------------------------------------------------------------ host utilities (no GPU)

def playouts(seed: int, n_games: int, max_plies: int, threads: int = 0):
    """Random legal playouts from the start position: (boards, moves, game_start)."""
    L = lib()
    threads = threads or min(os.cpu_count() or 1, 32)
    cap = n_games * (max_plies + 1)
    boards = np.zeros(cap, dtype=BOARD_DTYPE)
    moves = np.zeros(cap, dtype=np.uint16)
    starts = np.zeros(n_games + 1, dtype=np.uint32)
    n = L.sp_host_playouts(seed, n_games, max_plies, threads, boards.ctypes.data, moves.ctypes.data, starts.ctypes.data)
    return boards[:n].copy(), moves[:n].copy(), starts


def net_payload(net_image) -> np.ndarray:
    """Logical payload bytes of a network image, decompressing a zstd-flagged one (sp_host_net_payload)."""
    image = np.ascontiguousarray(net_image, dtype=np.uint8)
    out = np.empty(89_381_920, dtype=np.uint8)
    n = lib().sp_host_net_payload(image.ctypes.data, image.size, out.ctypes.data, out.size)
    if n < 0:
        raise NnueError(SP_ERR_BAD_NETWORK, "not a loadable network image")
    return out[:n]


def board_from_fen(fen: str) -> np.ndarray:
    out = np.zeros(1, dtype=BOARD_DTYPE)
    rc = lib().sp_host_board_from_fen(fen.encode(), out.ctypes.data)
    if rc:
        raise NnueError(rc, f"bad fen: {fen}")
    return out


def board_from_dfrc(index: int) -> np.ndarray:
    out = np.zeros(1, dtype=BOARD_DTYPE)
    rc = lib().sp_host_board_from_dfrc(int(index), out.ctypes.data)
    if rc:
        raise NnueError(rc, "bad DFRC index")
    return out


def board_to_fen(board) -> str:
    board = _boards(board).reshape(1)
    buf = C.create_string_buffer(128)
    rc = lib().sp_host_board_to_fen(board.ctypes.data, buf, 128)
    if rc:
        raise NnueError(rc, "bad board")
    return buf.value.decode()


def legal_moves(board) -> np.ndarray:
    board = _boards(board).reshape(1)
    out = np.zeros(256, dtype=np.uint16)
    n = lib().sp_host_legal_moves(board.ctypes.data, out.ctypes.data)
    if n < 0:
        raise NnueError(SP_ERR_BAD_BOARD, "bad board")
    return out[:n].copy()


def host_adjust(boards, raw, params: "AdjustParams", correction=None) -> np.ndarray:
    """adjustStatic + adjustEval through the C++ mirror's per-position form, on the host (sp_host_adjust)."""
    boards = _boards(boards)
    raw = np.ascontiguousarray(raw, dtype=np.int32)
    corr = None if correction is None else np.ascontiguousarray(correction, dtype=np.int32)
    out = np.empty(len(boards), dtype=np.int32)
    rc = lib().sp_host_adjust(boards.ctypes.data, raw.ctypes.data, None if corr is None else corr.ctypes.data, len(boards), C.byref(params), out.ctypes.data)
    if rc:
        raise NnueError(rc, "sp_host_adjust failed")
    return out


def in_check(board) -> bool:
    rc = lib().sp_host_in_check(_boards(board).ctypes.data)
    if rc < 0:
        raise ValueError("malformed board record")
    return bool(rc)


def apply_move(board, move: int) -> np.ndarray:
    board = _boards(board).reshape(1)
    out = np.zeros(1, dtype=BOARD_DTYPE)
    rc = lib().sp_host_apply_move(board.ctypes.data, int(move), out.ctypes.data)
    if rc:
        raise NnueError(rc, "bad board")
    return out


def features(board, perspective: int, kind: int) -> np.ndarray:
    """Feature indices from the shared host/device indexing code. kind 0 = PSQ, 1 = threat + pawn pair."""
    board = _boards(board).reshape(1)
    out = np.zeros(512, dtype=np.uint32)
    n = lib().sp_host_features(board.ctypes.data, perspective, kind, out.ctypes.data)
    if n < 0:
        raise NnueError(SP_ERR_BAD_BOARD, "bad board")
    return out[:n].copy()


def feature_counts(boards, threads: int = 0) -> dict:
    """Rows a full refresh of `boards` must read, summed over both perspectives."""
    boards = _boards(boards)
    out = np.zeros(3, dtype=np.uint64)
    rc = lib().sp_host_feature_counts(boards.ctypes.data, boards.size, threads or min(os.cpu_count() or 1, 32), out.ctypes.data)
    if rc:
        raise NnueError(rc, "bad board")
    return {"psq_rows": int(out[0]), "threat_rows": int(out[1]), "pawn_pair_rows": int(out[2])}


def playout_stats(boards, game_start, threads: int = 0) -> dict:
    """Rows the incremental walker must read for a playout stream (see sp_host_playout_stats)."""
    boards = _boards(boards)
    game_start = np.ascontiguousarray(game_start, dtype=np.uint32)
    out = np.zeros(6, dtype=np.uint64)
    rc = lib().sp_host_playout_stats(boards.ctypes.data, game_start.ctypes.data, game_start.size - 1,
                                     threads or min(os.cpu_count() or 1, 32), out.ctypes.data)
    if rc:
        raise NnueError(rc, "bad board")
    keys = ["psq_delta_rows", "threat_delta_rows", "rebuild_psq_rows", "rebuild_threat_rows", "updated_perspectives", "rebuilt_perspectives"]
    return {k: int(v) for k, v in zip(keys, out)}


def feature_delta(before, after, perspective: int):
    """(needs_refresh, psq_add, psq_sub, thr_add, thr_sub) from the shared delta generator."""
    before = _boards(before).reshape(1)
    after = _boards(after).reshape(1)
    bufs = [np.zeros(512, dtype=np.uint32) for _ in range(4)]
    counts = [C.c_int(0) for _ in range(4)]
    args = []
    for b, c in zip(bufs, counts):
        args += [b.ctypes.data, C.cast(C.byref(c), _vp)]
    rc = lib().sp_host_feature_delta(before.ctypes.data, after.ctypes.data, perspective, *args)
    if rc < 0:
        raise NnueError(SP_ERR_BAD_BOARD, "bad board")
    return (rc == 1, *[b[: c.value].copy() for b, c in zip(bufs, counts)])


class SelfplayParams(C.Structure):
    """SpSelfplayParams (include/sp_nnue.h)."""

    _fields_ = [
        ("concurrency", C.c_uint32),
        ("total_games", C.c_uint32),
        ("threads", C.c_uint32),
        ("depth", C.c_uint32),
        ("nodes_per_move", C.c_uint32),
        ("max_plies", C.c_uint32),
        ("seed", C.c_uint64),
        ("dfrc", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class SelfplayStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("games", "positions", "nodes", "evals", "batches", "searches")]


def selfplay(net_image, device: int = 0, *, concurrency: int = 1024, total_games: int = 1024, threads: int = 1, depth: int = 3,
             nodes_per_move: int = 5000, max_plies: int = 300, seed: int = 42, capacity: int | None = None, resident: bool = False, dfrc: bool = False):
    """Batched self-play (sp_selfplay_run; resident=True: sp_selfplay_run_gpu, the games' searches run on the device too).
    Returns (viriformat bytes as uint8 array, stats dict)."""
    img = np.ascontiguousarray(net_image, dtype=np.uint8)
    p = SelfplayParams(concurrency, total_games, threads, depth, nodes_per_move, max_plies, seed, int(dfrc), 0)
    st = SelfplayStats()
    cap = capacity if capacity is not None else total_games * (32 + 4 * (max_plies + 2))
    out = np.empty(cap, dtype=np.uint8)
    out_len = C.c_size_t(0)
    run = lib().sp_selfplay_run_gpu if resident else lib().sp_selfplay_run
    rc = run(img.ctypes.data, img.size, device, C.byref(p), C.byref(st), out.ctypes.data, cap, C.byref(out_len))
    if rc != SP_OK:
        raise NnueError(rc, (lib().sp_nnue_last_error(None) or b"").decode())
    return out[: out_len.value].copy(), {k: int(getattr(st, k)) for k, _ in SelfplayStats._fields_}


def parse_viriformat(data):
    """viriformat bytes -> list of (start board record, moves uint16[], scores int16[]) (src/datagen/viriformat.cpp:51-62)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    games, at = [], 0
    while at < len(data):
        start = data[at : at + 32].view(BOARD_DTYPE)[0].copy()
        at += 32
        pairs = data[at:].view("<u2").reshape(-1, 2)
        end = int(np.flatnonzero((pairs[:, 0] == 0) & (pairs[:, 1] == 0))[0])
        games.append((start, pairs[:end, 0].copy(), pairs[:end, 1].copy().view("<i2")))
        at += 4 * (end + 1)
    return games


def viri_to_move(board, viri: int) -> int:
    """viriformat move (from | to << 6 | promo << 12 | type flags) -> this library's SpMove, by matching the legal moves."""
    flags = {0x0000: 0, 0xC000: 1, 0x8000: 2, 0x4000: 3}[viri & 0xC000]
    frm, to, promo = viri & 63, (viri >> 6) & 63, (viri >> 12) & 3
    want = frm << 10 | to << 4 | (promo << 2 if flags == 1 else 0) | flags
    moves = legal_moves(board)
    if want not in moves:
        raise ValueError(f"viriformat move {viri:#06x} is not legal in {board_to_fen(board)}")
    return want


def viriformat(start, moves, scores, outcome: int) -> np.ndarray:
    """One game -> viriformat bytes through the driver's writer (sp_host_viriformat)."""
    start = _boards(start).reshape(1)
    moves = np.ascontiguousarray(moves, dtype=np.uint16)
    scores = np.ascontiguousarray(scores, dtype=np.int16)
    out = np.empty(32 + 4 * (len(moves) + 1), dtype=np.uint8)
    n = lib().sp_host_viriformat(start.ctypes.data, moves.ctypes.data, scores.ctypes.data, len(moves), outcome, out.ctypes.data, out.size)
    if n < 0:
        raise ValueError("sp_host_viriformat failed")
    return out[:n]


def normalize_score(board, score: int):
    """(classical material, wdl-normalised score) of a position (sp_host_normalize_score)."""
    material, norm = C.c_int32(), C.c_int32()
    rc = lib().sp_host_normalize_score(_boards(board).ctypes.data, int(score), C.byref(material), C.byref(norm))
    if rc:
        raise NnueError(rc, "bad board")
    return material.value, norm.value


def wdl_model(board, pov_score: int):
    """(win, loss) per mille of a side-to-move score (sp_host_wdl_model)."""
    win, loss = C.c_int32(), C.c_int32()
    rc = lib().sp_host_wdl_model(_boards(board).ctypes.data, int(pov_score), C.byref(win), C.byref(loss))
    if rc:
        raise NnueError(rc, "bad board")
    return win.value, loss.value
