"""MCTSTree with the reference's call surface (mcts/tree.py:26-519) on top of the CUDA engine.

    tree = MCTSTree(network, tree_size=65536, batch_size=1, cgos_mode=False)
    pos = tree.generate_move_with_sequential_halving(board, color, time_manager, never_resign)   # tree.py:318
    pos = tree.search_best_move(board, color, time_manager, analysis_query)                      # tree.py:57

`board` is anything with the reference GoBoard's get_move_history() / get_handicap_history() / get_board_size() /
get_komi(); it is not modified.  The whole search -- expansion, feature planes, DualNet, selection, backup -- runs on
the device for the single game of this tree; batches of games go through tamago_b200.selfplay instead.
"""
import json
import select
import sys
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

    def _run_puct(self, board, color, visits, strict):
        e = self._setup_position(board, color, visits, False)
        self.to_move = color
        self._board_size = board.get_board_size()
        return self._finish(e, e.genmove(mode=MODE_PUCT, visits=visits, strict=strict, play=False))

    @staticmethod
    def _is_strict(time_manager):
        return bool(getattr(time_manager, "is_strict", lambda: getattr(time_manager.mode, "name", "") == "STRICT_PLAYOUT")())

    def _write_analysis(self, board, analysis_query):
        mode = analysis_query.get("mode", "lz")
        sys.stdout.write(self.get_root().get_analysis(board, mode, self.get_pv_lists))
        sys.stdout.flush()

    def search_best_move(self, board, color, time_manager, analysis_query=None):
        """mcts/tree.py:57-105.  A non-empty analysis_query makes the search report like MCTSTree.search does (tree.py:155-174);
        the whole visit budget runs on the device in one call, so the report is written once, from the final tree."""
        visits = time_manager.get_num_visits_threshold(color)
        if hasattr(time_manager, "start_timer"):
            time_manager.start_timer()                                    # tree.py:71
        start = time.time()
        pos = self._run_puct(board, color, visits, self._is_strict(time_manager))
        root = self.get_root()
        if analysis_query and root.get_num_children() > 1:
            self._write_analysis(board, analysis_query)
        if root.get_num_children() > 1:                                   # tree.py:76-77 returns before any bookkeeping
            search_time = max(time.time() - start, 1e-9)
            if hasattr(time_manager, "set_search_speed"):
                time_manager.set_search_speed(int(root.node_visits), search_time)        # tree.py:94
            if hasattr(time_manager, "substract_consumption_time"):
                time_manager.substract_consumption_time(color, search_time)              # tree.py:95: TIME_CONTROL budgets shrink
        return pos

    def search(self, board, color, time_manager, analysis_query):
        """mcts/tree.py:130-174: run the PUCT budget of time_manager from `board` and, when analysis_query is not empty,
        write the lz / cgos analysis of the root to stdout."""
        self._run_puct(board, color, time_manager.get_num_visits_threshold(color), self._is_strict(time_manager))
        if analysis_query:
            self._write_analysis(board, analysis_query)

    def ponder(self, board, color, analysis_query):
        """mcts/tree.py:108-127: search without a visit limit until input arrives on stdin (gtp/client.py lz-analyze).  The
        device search runs fixed budgets, so pondering re-searches with a doubling budget (every round restarts the tree)
        and reports after each round.  It stops when stdin has input, when the budget reaches analysis_query["max_visits"]
        (default: the visits the device tree can hold, 65536 like MCTS_TREE_SIZE), or -- when the caller did not ask for
        stdin polling with the "ponder" flag -- after ONE round of analysis_query.get("visits", 1024) visits."""
        cap = int(analysis_query.get("max_visits", 65536))
        if not analysis_query.get("ponder", False):
            self._run_puct(board, color, min(cap, int(analysis_query.get("visits", 1024))), True)
            if analysis_query:
                self._write_analysis(board, analysis_query)
            return
        visits = min(cap, 256)
        self._get_engine(board, cap)                                      # one engine sized for the largest round
        while True:
            self._run_puct(board, color, visits, True)
            self._write_analysis(board, analysis_query)
            if visits >= cap:
                break
            try:
                rlist, _, _ = select.select([sys.stdin], [], [], 0)
            except (ValueError, OSError):
                rlist = [True]
            if rlist:
                break
            visits = min(cap, visits * 2)

    def search_with_callback(self, board, color, callback):
        """mcts/tree.py:177-196 (used by animation/ only): single-descent stepping is not exposed by the device engine."""
        raise NotImplementedError("search_with_callback: per-descent stepping is host-driven in the reference (animation.py); "
                                  "the device search runs whole budgets")

    def get_root(self):
        e, res = self._last
        k = int(res["num_children"][0])
        return MCTSNodeView(e.node(0, 0), improved=res["improved"][0, :k].copy(), max_actions=e.A)

    def _node_at(self, index):
        e, _ = self._last
        return MCTSNodeView(e.node(0, index), max_actions=e.A)

    @property
    def node(self):
        return [self._node_at(i) for i in range(self.num_nodes)]

    def to_dict(self):
        """mcts/tree.py:489-506."""
        return {"node": [nd.to_dict() for nd in self.node], "num_nodes": int(self.num_nodes), "root": 0,
                "current_root": int(self.current_root), "batch_size": int(self.batch_size), "cgos_mode": bool(self.cgos_mode),
                "to_move": "black" if color_value(self.to_move) == 1 else "white"}

    def dump_to_json(self, board, superko):
        """mcts/tree.py:476-486 + mcts/dump.py:10-33 (tamago-dump_tree)."""
        from ..program import PROGRAM_NAME, VERSION, PROTOCOL_VERSION
        state = {"dump_version": 2, "tree": self.to_dict(), "board_size": board.get_board_size(), "komi": board.get_komi(),
                 "move_history": [("black" if color_value(c) == 1 else "white", int(p)) for (c, p, *_r) in board.get_move_history()],
                 "handicap_history": [int(p) for p in board.get_handicap_history()], "superko": superko,
                 "name": PROGRAM_NAME, "version": VERSION, "protocol_version": PROTOCOL_VERSION}
        return json.dumps(state)

    def get_pv_lists(self, root, coord):
        """mcts/tree.py:432-449: principal variation below every visited root child."""
        pv = {}
        for i in range(root.get_num_children()):
            if root.children_visits[i] > 0:
                seq = self.get_best_move_sequence([root.get_child_move(i)], root.get_child_index(i))
                pv[coord.convert_to_gtp_format(root.get_child_move(i))] = [coord.convert_to_gtp_format(p) for p in seq]
        return pv

    def get_best_move_sequence(self, pv_list, index):
        """mcts/tree.py:451-473: follow the most visited child (first index on ties) until an unvisited or unexpanded node."""
        while index >= 0:                      # the reference indexes node[-1] for NOT_EXPANDED: an unused node, visits 0
            nd = self._node_at(index)
            if nd.node_visits == 0:
                break
            b = nd.get_best_move_index()
            pv_list.append(nd.get_child_move(b))
            index = nd.get_child_index(b)
        return pv_list
