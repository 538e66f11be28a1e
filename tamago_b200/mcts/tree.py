"""MCTSTree with the reference's call surface (mcts/tree.py:26-519) on top of the CUDA engine.

    tree = MCTSTree(network, tree_size=65536, batch_size=1, cgos_mode=False)
    pos = tree.generate_move_with_sequential_halving(board, color, time_manager, never_resign)   # tree.py:318
    pos = tree.search_best_move(board, color, time_manager, analysis_query)                      # tree.py:57

`board` is anything with the reference GoBoard's get_move_history() / get_handicap_history() / get_board_size() /
get_komi(); it is not modified.  The whole search -- expansion, feature planes, DualNet, selection, backup -- runs on
the device for the single game of this tree; batches of games go through tamago_b200.selfplay instead.
"""
import time

import numpy as np

from ..board.constant import PASS, RESIGN
from ..board.stone import color_value
from ..engine import Engine, MODE_SH, MODE_PUCT, EVAL_DUALNET_TC
from .constant import MCTS_TREE_SIZE, NN_BATCH_SIZE
from .node import MCTSNodeView


class MCTSTree:
    def __init__(self, network, tree_size=MCTS_TREE_SIZE, batch_size=NN_BATCH_SIZE, cgos_mode=False, seed=0):
        self.network = network
        self.tree_size = tree_size
        self.batch_size = batch_size
        self.cgos_mode = cgos_mode
        self.seed = seed
        self.current_root = 0
        self.num_nodes = 0
        self.to_move = None
        self._engine = None
        self._key = None
        self._last = None
        self._game_counter = 0

    # -- engine management ------------------------------------------------------------------------
    def _get_engine(self, board, visits):
        size, komi = board.get_board_size(), board.get_komi()
        superko = bool(getattr(board, "check_superko", False))
        key = (size, komi, superko, self.batch_size)
        if self._engine is None or self._key != key or self._max_visits < visits:
            if self._engine is not None:
                self._engine.close()
            self._max_visits = max(visits, 16)
            self._engine = Engine(board_size=size, games=1, max_visits=self._max_visits, komi=komi, superko=superko,
                                  batch_size=self.batch_size, evaluator=getattr(self.network, "evaluator", EVAL_DUALNET_TC),
                                  cgos_mode=self.cgos_mode, seed=self.seed, device=getattr(self.network, "device_index", 0),
                                  net_blocks=getattr(self.network, "blocks", 6))
            if getattr(self.network, "state_dict_np", None) is not None:
                self._engine.load_state_dict(self.network.state_dict_np)
            zob = getattr(board, "zobrist_table", None)
            if zob is not None:
                self._engine.set_zobrist(zob)
            self._key = key
        return self._engine

    def _setup_position(self, board, color, visits, never_resign):
        e = self._get_engine(board, visits)
        self._game_counter += 1
        e.reset(game_ids=[self._game_counter], never_resign=[1 if never_resign else 0])
        hist = [(color_value(c), int(p)) for (c, p, *_rest) in board.get_move_history()]
        handicap = [int(p) for p in board.get_handicap_history()] if hasattr(board, "get_handicap_history") else []
        seq = [(1, p) for p in handicap] + hist
        if seq:
            mv = np.array([[p for _, p in seq]], np.int16)
            col = np.array([[c for c, _ in seq]], np.uint8)
            e.play(mv, colors=col)
        e.set_to_move(color_value(color))
        return e

    def _finish(self, e, res):
        if res["error"][0]:
            raise RuntimeError(f"device search error flags {int(res['error'][0])} (history / depth / node pool overflow)")
        self.num_nodes = e.tree_size(0)
        self._last = (e, res)
        return int(res["move"][0])

    # -- reference surface ------------------------------------------------------------------------
    def generate_move_with_sequential_halving(self, board, color, time_manager, never_resign):
        visits = time_manager.get_num_visits_threshold(color)
        e = self._setup_position(board, color, visits, never_resign)
        self.to_move = color
        return self._finish(e, e.genmove(mode=MODE_SH, visits=visits, play=False))

    def search_best_move(self, board, color, time_manager, analysis_query=None):
        visits = time_manager.get_num_visits_threshold(color)
        strict = bool(getattr(time_manager, "is_strict", lambda: getattr(time_manager.mode, "name", "") == "STRICT_PLAYOUT")())
        e = self._setup_position(board, color, visits, False)
        self.to_move = color
        start = time.time()
        pos = self._finish(e, e.genmove(mode=MODE_PUCT, visits=visits, strict=strict, play=False))
        root = self.get_root()
        if hasattr(time_manager, "set_search_speed"):
            time_manager.set_search_speed(int(root.node_visits), time.time() - start)
        return pos

    def get_root(self):
        e, res = self._last
        k = int(res["num_children"][0])
        return MCTSNodeView(e.node(0, 0), improved=res["improved"][0, :k].copy())

    @property
    def node(self):
        e, _ = self._last
        return [MCTSNodeView(e.node(0, i)) for i in range(self.num_nodes)]

    def get_pv_lists(self, root, coord):
        """mcts/tree.py:432-473: principal variation below every root child that has been expanded."""
        e, _ = self._last
        pv = {}
        for i in range(root.get_num_children()):
            if root.children_visits[i] == 0:
                continue
            seq, idx = [coord.convert_to_gtp_format(root.get_child_move(i))], root.get_child_index(i)
            while idx >= 0:
                nd = MCTSNodeView(e.node(0, idx))
                if nd.num_children == 0 or nd.children_visits.max() == 0:
                    break
                b = nd.get_best_move_index()
                seq.append(coord.convert_to_gtp_format(nd.get_child_move(b)))
                idx = nd.get_child_index(b)
            pv[coord.convert_to_gtp_format(root.get_child_move(i))] = seq
        return pv
