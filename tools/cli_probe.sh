#!/bin/bash
# usage: tools/cli_probe.sh WxH frames  -- writes stills once, runs the CLI with -v in a few configurations
set -e
python - "$1" "$2" <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from toucan_b200 import render, workloads
w, h = [int(v) for v in sys.argv[1].split("x")]
frames = int(sys.argv[2])
d = "/tmp/cli_probe"; os.makedirs(d + "/stills", exist_ok=True)
rng = np.random.default_rng(5)
for t in range(10):
    for k in range(2):
        render.write_exr(os.path.join(d, f"stills/t{t}_k{k}.exr"), rng.random((h, w, 4), dtype=np.float32), zip=False)
for n in (frames, 2 * frames):
    text = json.dumps(workloads.stack_timeline(tracks=10, frames=n))
    for t in range(10):
        for k in range(2):
            text = text.replace(workloads.still_url(t, k), f"stills/t{t}_k{k}.exr")
    open(d + f"/stack_{n}.otio", "w").write(text)
PY
EXE=toucan_b200/lib/toucan-render
F=$2
echo "== warm"; $EXE /tmp/cli_probe/stack_$F.otio - -y4m 422 -v > /dev/null
echo "== N"; $EXE /tmp/cli_probe/stack_$F.otio - -y4m 422 -v > /dev/null
echo "== 2N"; $EXE /tmp/cli_probe/stack_$((2*F)).otio - -y4m 422 -v > /dev/null
echo "== 2N again"; $EXE /tmp/cli_probe/stack_$((2*F)).otio - -y4m 422 -v > /dev/null
rm -rf /tmp/cli_probe
