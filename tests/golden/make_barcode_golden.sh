#!/bin/bash
# Regenerates tests/golden/barcode_splitter/ by running the REFERENCE script (Perl) in this container:
#   /root/reference/scripts/fastx_barcode_splitter.pl on the reference's own fixture
#   (galaxy/test-data/fastx_barcode_splitter1.fastq + fastx_barcode_splitter1.txt, galaxy/tools/fastx_toolkit/fastx_barcode_splitter.xml:22-31)
# plus a FASTA form of the same reads.  Each case directory holds the script's output files and its summary (stdout).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/barcode_splitter
S=/root/reference/scripts/fastx_barcode_splitter.pl
T=/root/reference/galaxy/test-data
rm -rf "$OUT"; mkdir -p "$OUT"
cp $T/fastx_barcode_splitter1.fastq "$OUT/in.fastq"
cp $T/fastx_barcode_splitter1.txt "$OUT/barcodes.txt"
awk 'NR%4==1{print ">" substr($0,2)} NR%4==2{print}' "$OUT/in.fastq" > "$OUT/in.fasta"
run() {  # name input flags...
  name=$1; input=$2; shift 2
  mkdir -p "$OUT/$name"
  (cd "$OUT/$name" && perl $S --bcfile ../barcodes.txt --prefix out_ --suffix .txt "$@" < "../$input" > summary.txt)
  echo "$*" > "$OUT/$name/flags.txt"
}
run bol_mm2 in.fastq --bol --mismatches 2          # the reference's own test case
run bol_exact in.fastq --bol --exact
run eol_mm1 in.fastq --eol
run bol_partial2 in.fastq --bol --mismatches 2 --partial 2
run eol_partial1 in.fasta --eol --mismatches 1 --partial 1
run bol_fasta in.fasta --bol --mismatches 3
