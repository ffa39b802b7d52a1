#ifndef FXH_USAGE_H
#define FXH_USAGE_H
extern const char *const fxh_usage_fastq_masker;
extern const char *const fxh_usage_fastq_quality_filter;
extern const char *const fxh_usage_fastq_to_fasta;
extern const char *const fxh_usage_fastq_quality_trimmer;
extern const char *const fxh_usage_fastx_artifacts_filter;
extern const char *const fxh_usage_fastx_clipper;
extern const char *const fxh_usage_fastx_collapser;
extern const char *const fxh_usage_fastx_quality_stats;
extern const char *const fxh_usage_fastx_reverse_complement;
extern const char *const fxh_usage_fastx_trimmer;
#endif
