#!/usr/bin/env python
"""Regenerates tests/golden/pipeline.json: digests of the reference binaries (oracle/_ref) run in a shell pipe over the
seeded synthetic input of tests/test_pipeline_oracle.py.  Run from the repo root in the container that has /root/reference."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import test_pipeline_oracle as T  # noqa: E402

out = {}
seq, qual, L = T.synth_input()
with tempfile.TemporaryDirectory() as d:
    fq = os.path.join(d, "in.fq")
    H.write_fastq(fq, seq, qual, None, L)
    for name, (cmds, _) in sorted(T.PIPELINES.items()):
        data = open(fq, "rb").read()
        for cmd in cmds:
            data = subprocess.run([H.ref_tool(cmd[0])] + cmd[1:], input=data, stdout=subprocess.PIPE, check=True).stdout
        out[name] = {"sha256": hashlib.sha256(data).hexdigest(), "unique_sequences": data.count(b"\n") // 2,
                     "commands": [" ".join(c) for c in cmds]}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "pipeline.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
