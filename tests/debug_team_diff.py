"""TEST INFRASTRUCTURE (not collected by pytest): run one team-schedule search against the C oracle and report where the
trees differ.  Lives under tests/ because it imports oracle/.  python tests/debug_team_diff.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import liboracle as lo
from oracle import network_np as nn
from cumind_b200 import _native, params as P
from cumind_b200.network import CuMindNetwork
from cumind_b200.mcts import MCTS

class Cfg:
    def __init__(self, S, A, c_puct=1.25, alpha=0.3, frac=0.25):
        self.num_simulations, self.action_space_size, self.c_puct = S, A, c_puct
        self.dirichlet_alpha, self.exploration_fraction = alpha, frac

def main(A, H, S, B, geom):
    blob = nn.flatten_params(nn.random_params((4,), A, H, 2, 32, 11 + A, bias_scale=0.05), (4,))
    rng = np.random.default_rng(A * 1000 + H)
    obs = rng.standard_normal((B, 4)).astype(np.float32)
    desc = lo.make_desc((4,), A, H, 2)
    h, _, _ = lo.initial_inference(desc, blob, obs)
    noise = rng.dirichlet([0.3] * A, size=B)
    o = lo.search(desc, blob, h, S, A, 1.25, 0.25, noise)
    net = CuMindNetwork((4,), A, H, 2, 32, params=P.unflatten(blob, (4,), A, H, 2, 32))
    for rep in range(3):
        m = MCTS(net, Cfg(S, A))
        m.schedule = _native.SCHED_TEAM
        m.teams_per_cta, m.net_warps_per_team, m.trees_per_team, m.lanes_per_tree = geom
        probs, visits = m.search_batch(h, noise=noise)
        d = m.dump_trees()
        used = 1 + A * (S + 1)
        print("rep", rep, "plan", m.plan(B))
        for k in ("visit", "value_sum", "prior", "child_eidx", "exp_node", "hidden", "root_prior", "depth_sum"):
            a, b = d[k], o[k]
            if k in ("visit", "value_sum", "prior", "child_eidx"):
                a, b = a[:, :used], b[:, :used]
            elif k in ("exp_node", "hidden"):
                a, b = a[:, : S + 1], b[:, : S + 1]
            if a.dtype.kind == "f":
                u = {4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
                bad = a.view(u) != b.view(u)
            else:
                bad = a != b
            if bad.any():
                idx = np.argwhere(bad)
                trees = np.unique(idx[:, 0])
                print(f"  {k}: {bad.sum()} mismatches in {len(trees)} trees; first {idx[0]}; trees {trees[:10]} gpu={a[tuple(idx[0])]} ref={b[tuple(idx[0])]}")
                if k == "exp_node":
                    t = idx[0][0]
                    print("   gpu exp_node", a[t][:12], "\n   ref exp_node", b[t][:12])
            else:
                print(f"  {k}: ok")


if __name__ == "__main__":
    # usage: debug_team.py A,H,S,B[,teams,netwarps,trees,lanes] [more configs ...]  (run in order, one process)
    for arg in sys.argv[1:]:
        v = [int(x) for x in arg.split(",")]
        print("== config", v)
        main(v[0], v[1], v[2], v[3], tuple(v[4:8]) if len(v) >= 8 else (0, 0, 0, 0))
