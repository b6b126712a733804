"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

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

_vp = C.c_void_p


def _cpu_flags() -> set[str]:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_isa_available() -> list[str]:
    """ISA variants of the compiled reference this host can execute, best first."""
    flags = _cpu_flags()
    out = []
    need512 = {"avx512f", "avx512bw", "avx512vl", "avx512_vnni", "avx512vbmi", "avx512_vbmi2", "bmi2", "avx512dq", "avx512cd"}
    # -march=icelake-client also enables these; refuse the variant if the host lacks any of them
    need512 |= {"avx512_bitalg", "avx512_vpopcntdq", "gfni", "vaes", "vpclmulqdq", "sha_ni", "rdpid", "avx512ifma"}
    if need512 <= flags:
        out.append("avx512")
    if {"avx2", "bmi2", "fma", "movbe"} <= flags:
        out.append("avx2")
    return out


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


class Reference:
    """The reference's own CPU evaluation (oracle/_ref), via ref_shim.cpp."""

    def __init__(self, isa: str | None = None):
        avail = ref_isa_available()
        if isa is None:
            isa = next((v for v in avail if os.path.exists(self.path(v))), None)
        if isa is None or isa not in avail or not os.path.exists(self.path(isa)):
            raise FileNotFoundError(
                f"no runnable reference build (host supports {avail}); run `make -C oracle ref` where /root/reference exists"
            )
        self.isa = isa
        self.lib = C.CDLL(self.path(isa))
        L = self.lib
        L.spref_isa.restype = C.c_int
        L.spref_load_net.argtypes = [_vp, C.c_size_t]
        L.spref_eval_once.argtypes = [_vp, C.c_size_t, _vp]
        L.spref_time_eval_once.argtypes = [_vp, C.c_size_t, C.c_int, C.c_int, _vp]
        L.spref_time_eval_once.restype = C.c_double
        L.spref_playouts.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp, C.c_size_t]
        L.spref_playouts.restype = C.c_size_t
        L.spref_eval_playout.argtypes = [_vp, _vp, C.c_uint32, C.c_int, C.c_int, _vp]
        L.spref_time_playouts.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_int, C.c_int, _vp]
        L.spref_time_playouts.restype = C.c_double
        L.spref_psq_features.argtypes = [_vp, C.c_int, _vp]
        L.spref_threat_features.argtypes = [_vp, C.c_int, _vp]
        L.spref_threat_index.argtypes = [C.c_int] * 6
        L.spref_threat_index.restype = C.c_int32
        L.spref_legal_moves.argtypes = [_vp, _vp]
        L.spref_apply_move.argtypes = [_vp, C.c_uint16, _vp]
        L.spref_board_from_fen.argtypes = [C.c_char_p, _vp]
        L.spref_adjusted_eval.argtypes = [_vp, C.c_size_t, _vp, _vp, _vp, _vp]
        L.spref_viriformat.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_int, _vp, C.c_size_t]
        L.spref_viriformat.restype = C.c_long
        L.spref_normalize_score.argtypes = [_vp, C.c_int32, _vp, _vp]
        L.spref_board_from_dfrc.argtypes = [C.c_uint32, _vp]
        L.spref_wdl_model.argtypes = [_vp, C.c_int32, _vp, _vp]
        self._net = None

    @staticmethod
    def path(isa: str) -> str:
        return os.path.join(HERE, "_ref", f"libsp_ref_{isa}.so")

    @staticmethod
    def available() -> bool:
        return any(os.path.exists(Reference.path(v)) for v in ref_isa_available())

    def load_net(self, image: np.ndarray) -> None:
        image = np.ascontiguousarray(image, dtype=np.uint8)
        rc = self.lib.spref_load_net(_ptr(image), image.size)
        if rc:
            raise RuntimeError(f"spref_load_net failed ({rc})")

    def eval_once(self, boards: np.ndarray) -> np.ndarray:
        boards = np.ascontiguousarray(boards, dtype=BOARD_DTYPE)
        out = np.empty(boards.size, dtype=np.int32)
        rc = self.lib.spref_eval_once(_ptr(boards), boards.size, _ptr(out))
        if rc:
            raise RuntimeError(f"spref_eval_once failed ({rc})")
        return out

    def adjusted_eval(self, boards: np.ndarray, contempt=(0, 0), optimism=(0, 0)):
        """(raw evaluateOnce, staticEvalOnce + adjustEval<false>) from the reference's own functions."""
        boards = np.ascontiguousarray(boards, dtype=BOARD_DTYPE)
        raw = np.empty(boards.size, dtype=np.int32)
        out = np.empty(boards.size, dtype=np.int32)
        c = np.asarray(contempt, dtype=np.int32)
        o = np.asarray(optimism, dtype=np.int32)
        rc = self.lib.spref_adjusted_eval(_ptr(boards), boards.size, _ptr(c), _ptr(o), _ptr(raw), _ptr(out))
        if rc:
            raise RuntimeError(f"spref_adjusted_eval failed ({rc})")
        return raw, out

    def time_eval_once(self, boards: np.ndarray, threads: int, reps: int = 1):
        boards = np.ascontiguousarray(boards, dtype=BOARD_DTYPE)
        out = np.empty(boards.size, dtype=np.int32)
        secs = self.lib.spref_time_eval_once(_ptr(boards), boards.size, threads, reps, _ptr(out))
        if secs < 0:
            raise RuntimeError(f"spref_time_eval_once failed ({secs})")
        return secs, out

    def playouts(self, seed: int, n_games: int, max_plies: int, dfrc: bool = False, cap: int | None = None):
        cap = cap if cap is not None else n_games * (max_plies + 1)
        boards = np.zeros(cap, dtype=BOARD_DTYPE)
        moves = np.zeros(cap, dtype=np.uint16)
        starts = np.zeros(n_games + 1, dtype=np.uint32)
        n = self.lib.spref_playouts(seed, n_games, max_plies, int(dfrc), _ptr(boards), _ptr(moves), _ptr(starts), cap)
        return boards[:n].copy(), moves[:n].copy(), starts

    def eval_playout(self, start: np.ndarray, moves: np.ndarray, mode: int = 0, stride: int = 1) -> np.ndarray:
        start = np.ascontiguousarray(start, dtype=BOARD_DTYPE).reshape(1)
        moves = np.ascontiguousarray(moves, dtype=np.uint16)
        out = np.empty(moves.size + 1, dtype=np.int32)
        rc = self.lib.spref_eval_playout(_ptr(start), _ptr(moves), moves.size, mode, stride, _ptr(out))
        if rc:
            raise RuntimeError(f"spref_eval_playout failed ({rc})")
        return out

    def time_playouts(self, boards, moves, starts, threads: int, reps: int = 1):
        boards = np.ascontiguousarray(boards, dtype=BOARD_DTYPE)
        moves = np.ascontiguousarray(moves, dtype=np.uint16)
        starts = np.ascontiguousarray(starts, dtype=np.uint32)
        out = np.empty(boards.size, dtype=np.int32)
        secs = self.lib.spref_time_playouts(
            _ptr(boards), _ptr(moves), _ptr(starts), starts.size - 1, threads, reps, _ptr(out)
        )
        if secs < 0:
            raise RuntimeError(f"spref_time_playouts failed ({secs})")
        return secs, out

    def psq_features(self, board: np.ndarray, c: int) -> np.ndarray:
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        out = np.empty(32, dtype=np.uint32)
        n = self.lib.spref_psq_features(_ptr(board), c, _ptr(out))
        if n < 0:
            raise RuntimeError("spref_psq_features failed")
        return out[:n].copy()

    def threat_features(self, board: np.ndarray, c: int) -> np.ndarray:
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        out = np.empty(512, dtype=np.uint32)
        n = self.lib.spref_threat_features(_ptr(board), c, _ptr(out))
        if n < 0:
            raise RuntimeError("spref_threat_features failed")
        return out[:n].copy()

    def threat_index(self, c, king_sq, attacker, attacker_sq, attacked, attacked_sq) -> int:
        return self.lib.spref_threat_index(c, king_sq, attacker, attacker_sq, attacked, attacked_sq)

    def legal_moves(self, board: np.ndarray) -> np.ndarray:
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        out = np.empty(256, dtype=np.uint16)
        n = self.lib.spref_legal_moves(_ptr(board), _ptr(out))
        if n < 0:
            raise RuntimeError("spref_legal_moves failed")
        return out[:n].copy()

    def apply_move(self, board: np.ndarray, move: int) -> np.ndarray:
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        out = np.zeros(1, dtype=BOARD_DTYPE)
        rc = self.lib.spref_apply_move(_ptr(board), int(move), _ptr(out))
        if rc:
            raise RuntimeError("spref_apply_move failed")
        return out

    def wdl_model(self, board: np.ndarray, score: int):
        """wdl::wdlModel(score, pos.classicalMaterial()) -> (win, loss) per mille"""
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        win, loss = C.c_int32(), C.c_int32()
        if self.lib.spref_wdl_model(_ptr(board), int(score), C.byref(win), C.byref(loss)):
            raise RuntimeError("spref_wdl_model failed")
        return win.value, loss.value

    def board_from_dfrc(self, index: int) -> np.ndarray:
        out = np.zeros(1, dtype=BOARD_DTYPE)
        if self.lib.spref_board_from_dfrc(int(index), _ptr(out)):
            raise RuntimeError("spref_board_from_dfrc failed")
        return out

    def viriformat(self, start: np.ndarray, moves, scores, outcome: int) -> np.ndarray:
        """The reference's own Viriformat writer (src/datagen/viriformat.cpp) on one game."""
        start = np.ascontiguousarray(start, dtype=BOARD_DTYPE).reshape(1)
        moves = np.ascontiguousarray(moves, dtype=np.uint16)
        scores = np.ascontiguousarray(scores, dtype=np.int16)
        out = np.empty(32 + 4 * (len(moves) + 1), dtype=np.uint8)
        n = self.lib.spref_viriformat(_ptr(start), _ptr(moves), _ptr(scores), len(moves), outcome, _ptr(out), out.size)
        if n < 0:
            raise RuntimeError("spref_viriformat failed")
        return out[:n]

    def normalize_score(self, board: np.ndarray, score: int):
        """(pos.classicalMaterial(), wdl::normalizeScore<false>(score, material))"""
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        material, norm = C.c_int32(), C.c_int32()
        if self.lib.spref_normalize_score(_ptr(board), int(score), C.byref(material), C.byref(norm)):
            raise RuntimeError("spref_normalize_score failed")
        return material.value, norm.value

    def board_from_fen(self, fen: str) -> np.ndarray:
        out = np.zeros(1, dtype=BOARD_DTYPE)
        rc = self.lib.spref_board_from_fen(fen.encode(), _ptr(out))
        if rc:
            raise RuntimeError(f"bad fen: {fen}")
        return out


def build_c_oracle() -> str:
    """Compile oracle/nnue_oracle.c (gcc only) and return the path of the shared library."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    return os.path.join(HERE, "_build", "libsp_oracle.so")


class COracle:
    """The plain-C restatement (oracle/nnue_oracle.c)."""

    def __init__(self):
        self.lib = C.CDLL(build_c_oracle())
        L = self.lib
        L.spo_load_net.argtypes = [_vp, C.c_size_t]
        L.spo_eval_once.argtypes = [_vp, C.c_size_t, _vp]
        L.spo_time_eval_once.argtypes = [_vp, C.c_size_t, C.c_int, C.c_int, _vp]
        L.spo_time_eval_once.restype = C.c_double
        L.spo_psq_features.argtypes = [_vp, C.c_int, _vp]
        L.spo_threat_features.argtypes = [_vp, C.c_int, _vp]
        L.spo_threat_index.argtypes = [C.c_int] * 6
        L.spo_threat_index.restype = C.c_int32
        L.spo_accumulators.argtypes = [_vp, _vp, _vp]
        L.spo_ft_activations.argtypes = [_vp, _vp, _vp]
        L.spo_forward.argtypes = [_vp, C.c_int]
        L.spo_forward.restype = C.c_int32
        L.spo_forward_acc.argtypes = [_vp, _vp, C.c_int, C.c_int]
        L.spo_forward_acc.restype = C.c_int32

    def load_net(self, image: np.ndarray) -> None:
        image = np.ascontiguousarray(image, dtype=np.uint8)
        rc = self.lib.spo_load_net(_ptr(image), image.size)
        if rc:
            raise RuntimeError(f"spo_load_net failed ({rc})")

    def eval_once(self, boards: np.ndarray) -> np.ndarray:
        boards = np.ascontiguousarray(boards, dtype=BOARD_DTYPE)
        out = np.empty(boards.size, dtype=np.int32)
        rc = self.lib.spo_eval_once(_ptr(boards), boards.size, _ptr(out))
        if rc:
            raise RuntimeError(f"spo_eval_once failed ({rc})")
        return out

    def time_eval_once(self, boards: np.ndarray, threads: int, reps: int = 1):
        boards = np.ascontiguousarray(boards, dtype=BOARD_DTYPE)
        out = np.empty(boards.size, dtype=np.int32)
        secs = self.lib.spo_time_eval_once(_ptr(boards), boards.size, threads, reps, _ptr(out))
        if secs < 0:
            raise RuntimeError(f"spo_time_eval_once failed ({secs})")
        return secs, out

    def psq_features(self, board: np.ndarray, c: int) -> np.ndarray:
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        out = np.empty(32, dtype=np.uint32)
        n = self.lib.spo_psq_features(_ptr(board), c, _ptr(out))
        if n < 0:
            raise RuntimeError("spo_psq_features failed")
        return out[:n].copy()

    def threat_features(self, board: np.ndarray, c: int) -> np.ndarray:
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        out = np.empty(512, dtype=np.uint32)
        n = self.lib.spo_threat_features(_ptr(board), c, _ptr(out))
        if n < 0:
            raise RuntimeError("spo_threat_features failed")
        return out[:n].copy()

    def threat_index(self, c, king_sq, attacker, attacker_sq, attacked, attacked_sq) -> int:
        return self.lib.spo_threat_index(c, king_sq, attacker, attacker_sq, attacked, attacked_sq)

    def accumulators(self, board: np.ndarray):
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        psq = np.empty((2, 1024), dtype=np.int16)
        thr = np.empty((2, 1024), dtype=np.int16)
        rc = self.lib.spo_accumulators(_ptr(board), _ptr(psq), _ptr(thr))
        if rc:
            raise RuntimeError("spo_accumulators failed")
        return psq, thr

    def ft_activations(self, board: np.ndarray):
        board = np.ascontiguousarray(board, dtype=BOARD_DTYPE).reshape(1)
        out = np.empty(1024, dtype=np.uint8)
        bucket = C.c_int(0)
        rc = self.lib.spo_ft_activations(_ptr(board), _ptr(out), C.byref(bucket))
        if rc:
            raise RuntimeError("spo_ft_activations failed")
        return out, bucket.value

    def forward(self, ft: np.ndarray, bucket: int) -> int:
        ft = np.ascontiguousarray(ft, dtype=np.uint8)
        assert ft.size == 1024
        return self.lib.spo_forward(_ptr(ft), int(bucket))

    def forward_acc(self, psq: np.ndarray, thr: np.ndarray, stm: int, bucket: int) -> int:
        psq = np.ascontiguousarray(psq, dtype=np.int16)
        thr = np.ascontiguousarray(thr, dtype=np.int16)
        return self.lib.spo_forward_acc(_ptr(psq), _ptr(thr), int(stm), int(bucket))
