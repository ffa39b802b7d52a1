#!/bin/bash
# diagnostic: tools under tiny reader windows / text chunks
T=$(mktemp -d /tmp/fxg_diag.XXXX)
python - <<PY
import sys; sys.path.insert(0, "tests")
import numpy as np, helpers as H
seq, qual = H.synth_slab(H.SEED_BASE + 11, 40000, 100, H.WITH_N)
lens = H.ragged(seq, qual, np.random.default_rng(5), min_len=6)
H.write_fastq("$T/in.fq", seq, qual, lens, 100)
PY
for env in "FASTX_WINDOW_BYTES=70000 FASTX_CHUNK_BYTES=20000" "FASTX_WINDOW_BYTES=300000 FASTX_CHUNK_BYTES=100000" "FASTX_WINDOW_BYTES=70000 FASTX_TEXT_PATH=0 FASTX_BATCH_READS=777"; do
  for tool in "fastx_reverse_complement" "fastx_quality_stats" "fastx_clipper -a AGATCGGAAGAGC -l 10 -n -v"; do
    echo "== $env $tool"
    env $env ./bin/$tool -i $T/in.fq > $T/mine.out 2> $T/mine.err; echo "rc=$?"; head -c 600 $T/mine.err
    ./oracle/_ref/$tool -i $T/in.fq > $T/ref.out 2>$T/ref.err; cmp $T/mine.out $T/ref.out && echo same
  done
done
rm -rf $T
