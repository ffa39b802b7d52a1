/*
 * fxg_synth — write the deterministic synthetic workload (include/fxg_synth.h) as
 * FASTQ/FASTA text, so the reference CPU tools and the GPU tools can be fed identical input.
 *
 *   fxg_synth -n READS -l LEN [-k plain|n|adapter|dups] [-s SEED] [-f FIRST] [-T TOTAL] [-Q 33] [-a] [-o FILE]
 *     -a  FASTA instead of FASTQ          ids are "r<index>", line 3 is "+"
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fxg_synth.h"

int main(int argc, char **argv)
{
    long long n = 1000, first = 0, total = -1;
    int len = 100, kind = FXG_SYNTH_PLAIN, q_offset = 33, fasta = 0, opt;
    unsigned long long seed = FXG_SYNTH_SEED_BASE;
    const char *out = "-";
    while ((opt = getopt(argc, argv, "n:l:k:s:f:T:Q:ao:")) != -1) {
        switch (opt) {
        case 'n': n = atoll(optarg); break;
        case 'l': len = atoi(optarg); break;
        case 'k':
            if (!strcmp(optarg, "plain")) kind = FXG_SYNTH_PLAIN;
            else if (!strcmp(optarg, "n")) kind = FXG_SYNTH_WITH_N;
            else if (!strcmp(optarg, "adapter")) kind = FXG_SYNTH_ADAPTER;
            else if (!strcmp(optarg, "dups")) kind = FXG_SYNTH_DUPS;
            else { fprintf(stderr, "fxg_synth: unknown kind '%s'\n", optarg); return 1; }
            break;
        case 's': seed = strtoull(optarg, NULL, 10); break;
        case 'f': first = atoll(optarg); break;
        case 'T': total = atoll(optarg); break;
        case 'Q': q_offset = atoi(optarg); break;
        case 'a': fasta = 1; break;
        case 'o': out = optarg; break;
        default: fprintf(stderr, "usage: fxg_synth -n READS -l LEN [-k plain|n|adapter|dups] [-s SEED] [-f FIRST] [-T TOTAL] [-Q N] [-a] [-o FILE]\n"); return 1;
        }
    }
    if (total < 0) total = first + n;
    if (len <= 0 || len > 25000 || n < 0) { fprintf(stderr, "fxg_synth: bad geometry\n"); return 1; }
    FILE *f = strcmp(out, "-") ? fopen(out, "w") : stdout;
    if (!f) { perror(out); return 1; }
    static char iobuf[1 << 22];
    setvbuf(f, iobuf, _IOFBF, sizeof iobuf);
    char *line = (char *)malloc((size_t)2 * len + 64);
    for (long long i = first; i < first + n; i++) {
        unsigned long long r = fxg_synth_read_key(seed, (unsigned long long)i, kind, (unsigned long long)total);
        int p = sprintf(line, "%c%s%lld\n", fasta ? '>' : '@', "r", i);
        for (int k = 0; k < len; k++) line[p++] = (char)fxg_synth_base(r, k, len, kind);
        line[p++] = '\n';
        if (!fasta) {
            line[p++] = '+'; line[p++] = '\n';
            for (int k = 0; k < len; k++) line[p++] = (char)(fxg_synth_phred(r, k, len) + q_offset);
            line[p++] = '\n';
        }
        if (fwrite(line, 1, (size_t)p, f) != (size_t)p) { perror("write"); return 1; }
    }
    free(line);
    if (f != stdout) fclose(f); else fflush(f);
    return 0;
}
