"""Experiment: does kind::tf32 truncate or round its fp32 operands?  Runs the fused RGCN kernel with the splitter
writing hi = x & 0xFFFFE000 (default) and with the raw tile left in place as the hi operand (TFGNN_B200_DEBUG_SKIP=32),
in two processes (the knob is read once), and compares the outputs bitwise."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from tf2_gnn_b200.layers import MessagePassingInput, get_message_passing_class
rng = np.random.default_rng(0)
V, D, H, L = 20000, 256, 256, 4
adjs = [rng.integers(0, V, size=(60000, 2)).astype(np.int32) for _ in range(L)]
cls = get_message_passing_class("rgcn"); p = cls.get_default_hyperparameters(); p.update(hidden_dim=H, b200_path="fused_tc")
layer = cls(p); torch.manual_seed(0); layer.build(MessagePassingInput((None, D), tuple((None, 2) for _ in range(L))))
h = torch.from_numpy(rng.uniform(-1, 1, (V, D)).astype(np.float32)).cuda()
out = layer(MessagePassingInput(h, tuple(torch.from_numpy(a).cuda() for a in adjs)))
np.save(sys.argv[1], out.cpu().numpy())
''' % ROOT


def run(tag, env):
    path = f"/tmp/rawhi_{tag}.npy"
    subprocess.run([sys.executable, "-c", CHILD, path], check=True, env={**os.environ, **env})
    return np.load(path)


a = run("split", {})
b = run("raw", {"TFGNN_B200_DEBUG_SKIP": "32"})
print("bitwise equal:", np.array_equal(a, b), " max abs diff:", float(np.abs(a - b).max()), " max |out|:", float(np.abs(a).max()))
