#!/bin/bash
# usage: tools/cli_probe.sh WxH frames  -- writes stills, runs the CLI with -v in a few configurations
set -e
python - "$1" "$2" <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from toucan_b200 import render, workloads
w, h = [int(v) for v in sys.argv[1].split("x")]
frames = int(sys.argv[2])
d = "/tmp/cli_probe"; os.makedirs(d + "/stills", exist_ok=True)
text = json.dumps(workloads.stack_timeline(tracks=10, frames=frames))
rng = np.random.default_rng(5)
for t in range(10):
    for k in range(2):
        rel = f"stills/t{t}_k{k}.exr"
        render.write_exr(os.path.join(d, rel), rng.random((h, w, 4), dtype=np.float32), zip=False)
        text = text.replace(workloads.still_url(t, k), rel)
open(d + "/stack.otio", "w").write(text)
PY
EXE=toucan_b200/lib/toucan-render
echo "== y4m 422"; $EXE /tmp/cli_probe/stack.otio - -y4m 422 -v > /dev/null
echo "== raw"; $EXE /tmp/cli_probe/stack.otio - -v > /dev/null
echo "== no fusion"; TOUCAN_NO_FUSION=1 $EXE /tmp/cli_probe/stack.otio - -y4m 422 -v > /dev/null
rm -rf /tmp/cli_probe
