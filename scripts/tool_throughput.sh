#!/bin/bash
# file -> file throughput of the drop-in tools vs the reference binaries on the same synthetic FASTQ (one GPU box)
N=${1:-4000000}
T=$(mktemp -d /tmp/fxg_tt.XXXX)
./bin/fxg_synth -n $N -l 150 -k adapter -o $T/in.fq
ls -la $T/in.fq | awk '{print "input bytes:", $5}'
run() { local s=$(date +%s.%N); "$@" > /dev/null 2>$T/err; local e=$(date +%s.%N); echo "$s $e" | awk -v n=$N '{printf "%.2f s  %.2f Mreads/s\n", $2-$1, n/($2-$1)/1e6}'; }
for tool in fastq_quality_trimmer fastq_quality_filter fastx_reverse_complement fastx_clipper fastx_quality_stats fastx_collapser; do
  case $tool in
    fastq_quality_trimmer) A="-t 20 -l 20";; fastq_quality_filter) A="-q 20 -p 90";; fastx_clipper) A="-a AGATCGGAAGAGC -l 20";; *) A="";;
  esac
  echo -n "b200 $tool: "; run ./bin/$tool $A -i $T/in.fq -o $T/out_b200
  if [ "$tool" = "fastx_clipper" ]; then head -n $((N/10*4)) $T/in.fq > $T/in_small.fq; M=$((N/10)); else cp -l $T/in.fq $T/in_small.fq 2>/dev/null || cp $T/in.fq $T/in_small.fq; M=$N; fi
  s=$(date +%s.%N); ./oracle/_ref/$tool $A -i $T/in_small.fq -o $T/out_ref > /dev/null 2>&1; e=$(date +%s.%N)
  echo "$s $e" | awk -v n=$M -v t=$tool '{printf "ref  %s: %.2f s  %.3f Mreads/s (n=%d)\n", t, $2-$1, n/($2-$1)/1e6, n}'
  if [ "$M" = "$N" ]; then cmp $T/out_b200 $T/out_ref && echo "  identical output"; fi
  rm -f $T/in_small.fq
done
rm -rf $T
