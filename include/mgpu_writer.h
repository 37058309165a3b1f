/* mgpu_writer.h -- index writer (format v62) + the seeded synthetic corpus, as a library of its own (libmgpu_writer.so).
 *
 * Host-only: no CUDA, no query path.  Kept out of libmgpu.so so that processes which only need index files (the CPU
 * oracle's arm of bench.py, the tests' golden corpora) never map the GPU library.  SURVEY 8(f) rank 1: the byte layout of
 * CSphHitBuilder::cidxHit/cidxDone + IndexWriteHeader + CSphDictKeywords (src/sphinx.cpp:8297-8936, 19374-19700); it exists
 * because the reference's `indexer` cannot be built in this image. */
#ifndef MGPU_WRITER_H_
#define MGPU_WRITER_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgpu_build_doc_input {
	int32_t			n_docs;
	int32_t			n_fields;
	const char * const * field_names;
	int32_t			n_attrs;          /* uint32 attributes besides `id` */
	const char * const * attr_names;
	const int64_t *	docids;           /* [n_docs] ascending */
	const uint32_t *attrs;            /* [n_docs][n_attrs] */
	int32_t			n_keywords;
	const char * const * keywords;    /* dictionary forms */
	const int64_t *	field_tok_offsets;/* [n_docs*n_fields+1] into tok_* */
	const int32_t *	tok_keyword;      /* keyword index */
	const int32_t *	tok_pos;          /* 1-based position inside the field (gaps allowed) */
	int32_t			skiplist_block;   /* 0 -> 32 */
	int32_t			hit_format_inline;/* 1 = inline (default), 0 = plain */
	int32_t			dict_crc;         /* 0 = dict=keywords (default); 1 = dict=crc: entries keyed by sphFNV64 word ids
	                                     (CSphDiskDictTraits, src/sphinx.cpp:18263-18339) */
} mgpu_build_doc_input;
int				mgpu_build_index ( const char * path_prefix, const mgpu_build_doc_input * in, char * err, int errlen );

/* synthetic Zipfian corpus: docs [first_doc, first_doc+n_docs) of the seeded corpus are
 * written as a self-contained index with local rowids from 0 (a contiguous rowid-range shard). */
typedef struct mgpu_synth_params {
	uint64_t		seed;
	int64_t			first_doc;
	int64_t			n_docs;
	int32_t			vocab;            /* number of distinct terms, Zipf(s=1) */
	int32_t			title_min, title_max;
	int32_t			body_min, body_max;
	float			body_mu, body_sigma; /* lognormal */
	int32_t			threads;          /* 0 = all */
} mgpu_synth_params;
int				mgpu_build_synthetic ( const char * path_prefix, const mgpu_synth_params * p, char * err, int errlen );
/* token at (doc, field, pos0) of the synthetic corpus, and field length; lets query generators sample phrases */
int32_t			mgpu_synth_field_len ( const mgpu_synth_params * p, int64_t doc, int field );
int32_t			mgpu_synth_token ( const mgpu_synth_params * p, int64_t doc, int field, int pos0 );

int				mgpu_writer_abi_version ( void );

#ifdef __cplusplus
}
#endif
#endif /* MGPU_WRITER_H_ */
