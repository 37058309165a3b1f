/*
 * mgpu.h -- C ABI of the B200-native full-text query hot path.
 *
 * Drop-in boundary for Manticore Search 3.6.0's local-index search
 * (CSphIndex_VLN::MultiQuery, src/sphinx.cpp:15362; the seam is ParsedMultiQuery,
 * src/sphinx.cpp:15664-15943: sphCreateRanker + MatchExtended fill the caller's sorters).
 *
 * Every entry point below names the reference interface it replaces.  Plain C types only:
 * no C++ classes, no torch types, no exceptions cross this boundary.  There is NO CPU
 * fallback behind these calls: an operator or option the CUDA path does not implement
 * returns MGPU_E_UNSUPPORTED; a missing GPU returns MGPU_E_NO_DEVICE.
 */
#ifndef MGPU_H_
#define MGPU_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGPU_ABI_VERSION 1

/* ---- status codes (reference: bool return + tMeta.m_sError, src/sphinx.cpp:15690) ---- */
enum {
	MGPU_OK = 0,
	MGPU_E_IO = -1,           /* index files missing / unreadable */
	MGPU_E_FORMAT = -2,       /* not an index-format v57..v62 plain index (dict=keywords or dict=crc), or hitless / unsupported settings */
	MGPU_E_UNSUPPORTED = -3,  /* query uses an operator/ranker/sort the CUDA path does not implement */
	MGPU_E_BAD_QUERY = -4,    /* malformed tree (reference: sphCreateRanker returns nullptr, sphinxsearch.cpp:4377) */
	MGPU_E_NO_DEVICE = -5,    /* no CUDA device; the product never falls back to the CPU */
	MGPU_E_CUDA = -6,         /* CUDA runtime error, see mgpu_last_error() */
	MGPU_E_NOMEM = -7
};

/* ---- XQOperator_e (src/sphinxquery.h:43-62); only the listed values are accepted ---- */
enum {
	MGPU_OP_AND = 0,        /* SPH_QUERY_AND */
	MGPU_OP_OR = 1,         /* SPH_QUERY_OR */
	MGPU_OP_MAYBE = 2,      /* SPH_QUERY_MAYBE */
	MGPU_OP_NOT = 3,        /* SPH_QUERY_NOT: only as the one-child wrapper FixupNots leaves on the right side of an ANDNOT */
	MGPU_OP_ANDNOT = 4,     /* SPH_QUERY_ANDNOT */
	MGPU_OP_BEFORE = 5,     /* SPH_QUERY_BEFORE: a << b << c; on the GPU path when every child is a plain keyword */
	MGPU_OP_PHRASE = 6,     /* SPH_QUERY_PHRASE */
	MGPU_OP_PROXIMITY = 7,  /* SPH_QUERY_PROXIMITY, oparg = N of "..."~N */
	MGPU_OP_QUORUM = 8,     /* SPH_QUERY_QUORUM, oparg = threshold (absolute); one real quorum node per query on the GPU path */
	MGPU_OP_NEAR = 9,       /* SPH_QUERY_NEAR, oparg = distance; n-ary in the reference, two plain keywords on the GPU path */
	MGPU_OP_NOTNEAR = 10,   /* SPH_QUERY_NOTNEAR, oparg = distance; two children: must, not (plain keywords on the GPU path) */
	MGPU_OP_SENTENCE = 11,  /* SPH_QUERY_SENTENCE: mgpu_parse_query produces it; the evaluators answer MGPU_E_UNSUPPORTED (needs index_sp boundary hits) */
	MGPU_OP_PARAGRAPH = 12  /* SPH_QUERY_PARAGRAPH: likewise */
};

/* ---- ESphRankMode (src/sphinx.h:2388-2404) ---- */
enum {
	MGPU_RANK_PROXIMITY_BM25 = 0,
	MGPU_RANK_BM25 = 1,
	MGPU_RANK_NONE = 2,
	MGPU_RANK_WORDCOUNT = 3,
	MGPU_RANK_PROXIMITY = 4,   /* SPH_RANK_PROXIMITY (SPH01): RankerState_Proximity_fn<false,..> */
	MGPU_RANK_MATCHANY = 5,    /* SPH_RANK_MATCHANY (SPH02): RankerState_MatchAny_fn */
	MGPU_RANK_FIELDMASK = 6,   /* SPH_RANK_FIELDMASK: RankerState_Fieldmask_fn */
	MGPU_RANK_SPH04 = 7        /* SPH_RANK_SPH04: RankerState_ProximityBM25Exact_fn */
};

/* ---- sort key parts: ESphSortKeyPart (src/sortsetup.h:19-57) ---- */
enum {
	MGPU_KEYPART_ROWID = 0,
	MGPU_KEYPART_WEIGHT = 1,
	MGPU_KEYPART_INT = 2,     /* fixed-width integer row attribute (32 or 64 bits, dword aligned) */
	MGPU_KEYPART_FLOAT = 3    /* SPH_KEYPART_FLOAT (src/sphinxsort.cpp:4690-4696): a 32-bit row attribute compared as an IEEE float;
	                             -0 ties with +0 as in the reference; NaNs (no order in the reference) sort by their bit pattern */
};

/* ---- filters: the integer subset of ISphFilter::Eval (src/sphinxfilter.h:51) ---- */
enum {
	MGPU_FILTER_RANGE = 0,    /* min <= attr <= max  (SPH_FILTER_RANGE, closed bounds) */
	MGPU_FILTER_VALUES = 1    /* attr IN (values)    (SPH_FILTER_VALUES) */
};

/* XQKeyword_t (src/sphinxquery.h:21-39) */
typedef struct mgpu_xqkeyword {
	const char *	word;        /* m_sWord, dictionary form (tokenised, lower-cased by the caller's tokenizer) */
	int32_t			atom_pos;    /* m_iAtomPos, 1-based in-query position */
	float			boost;       /* m_fBoost (1.0f default) */
	uint8_t			field_start; /* m_bFieldStart: ^keyword (ExtTermPos_T, hit stage) */
	uint8_t			field_end;   /* m_bFieldEnd: keyword$ */
	uint8_t			excluded;    /* m_bExcluded */
	uint8_t			expanded;    /* m_bExpanded */
} mgpu_xqkeyword;

/* XQNode_t (src/sphinxquery.h:134-280), flattened: children and words are index ranges */
typedef struct mgpu_xqnode {
	int32_t			op;            /* MGPU_OP_* */
	int32_t			oparg;         /* m_iOpArg */
	int32_t			first_child;   /* index into mgpu_query.children[] */
	int32_t			n_children;
	int32_t			first_word;    /* index into mgpu_query.words[] */
	int32_t			n_words;
	uint32_t		field_mask;    /* XQLimitSpec_t::m_dFieldMask, fields 0..31; 0xFFFFFFFF = all */
	int32_t			field_max_pos; /* XQLimitSpec_t::m_iFieldMaxPos: @field[N] (0 = no limit) */
	uint8_t			not_weighted;  /* m_bNotWeighted */
	uint8_t			pad[3];
} mgpu_xqnode;

/* one key of CSphMatchComparatorState (src/sortsetup.h:19-57) */
typedef struct mgpu_sortkey {
	int32_t			kind;          /* MGPU_KEYPART_* */
	int32_t			attr;          /* attribute index in the index schema (MGPU_KEYPART_INT / MGPU_KEYPART_FLOAT) */
	int32_t			desc;          /* bit of m_uAttrDesc */
} mgpu_sortkey;

/* CSphFilterSettings subset (src/sphinx.h:2249-2320) */
typedef struct mgpu_filter {
	int32_t			kind;          /* MGPU_FILTER_* */
	int32_t			attr;          /* attribute index in the index schema */
	int64_t			min_value;     /* RANGE */
	int64_t			max_value;     /* RANGE */
	const int64_t *	values;        /* VALUES */
	int32_t			n_values;
	int32_t			exclude;       /* m_bExclude */
} mgpu_filter;

/* CSphQuery subset (src/sphinx.h:2586-2691) + CSphMultiQueryArgs (src/sphinx.h:2909-2925) */
typedef struct mgpu_query {
	const mgpu_xqnode *		nodes;
	int32_t					n_nodes;
	int32_t					root;           /* XQQuery_t::m_pRoot */
	const int32_t *			children;       /* flattened child index lists */
	int32_t					n_children;
	const mgpu_xqkeyword *	words;
	int32_t					n_words;

	int32_t					ranker;         /* m_eRanker, MGPU_RANK_* */
	const int32_t *			field_weights;  /* CSphQueryContext::m_dWeights after BindWeights (sphinx.cpp:13903); NULL = all 1 */
	int32_t					n_field_weights;

	const mgpu_sortkey *	sort_keys;      /* NULL/0 = SPH_SORT_RELEVANCE (weight desc, rowid asc) */
	int32_t					n_sort_keys;    /* <= 5; ties always broken by rowid asc (sphinxsort.cpp:4541-4790) */
	const mgpu_filter *		filters;
	int32_t					n_filters;

	int32_t					max_matches;    /* m_iMaxMatches (default 1000) -- the sorter size, not LIMIT */
	int32_t					index_weight;   /* iIndexWeight of MatchExtended (sphinx.cpp:12222); 0 -> 1 */
	uint8_t					plain_idf;      /* m_bPlainIDF */
	uint8_t					unnormalized_tfidf; /* !m_bNormalizedTFIDF */
	uint8_t					shard_of_global;    /* this index is one rowid-range shard and word_docs holds the statistics of the whole index:
	                                             order the keywords of multi-AND / phrase / quorum nodes (GetDocsCount, src/searchnode.cpp:2791)
	                                             and with them the fp32 TF*IDF additions by word_docs, as the unsharded index would
	                                             (SURVEY 8(e)); 0 = the reference's own per-index order */
	uint8_t					pad[1];

	/* CSphMultiQueryArgs::m_iTotalDocs / m_pLocalDocs (global IDF inputs for sharded indexes) */
	int64_t					total_docs;     /* 0 = use the index's own document count */
	const int64_t *			word_docs;      /* per words[] entry; NULL or <0 entries = use local dictionary docs */
} mgpu_query;

/* per-keyword statistics: CSphQueryResultMeta::AddStat (src/sphinxsearch.cpp:4365-4371) */
typedef struct mgpu_wordstat {
	int64_t			docs;
	int64_t			hits;
} mgpu_wordstat;

/* what ISphMatchSorter::Flatten would return (src/sphinxsort.cpp:627-641), caller-allocated */
typedef struct mgpu_result {
	int32_t			status;         /* MGPU_OK or error for this query (reference: m_iMultiplier=-1) */
	int32_t			n_matches;      /* <= max_matches */
	int64_t			total_found;    /* ISphMatchSorter::GetTotalCount() */
	uint32_t *		rowid;          /* [max_matches] best first */
	int32_t *		weight;         /* [max_matches] */
	int64_t *		docid;          /* [max_matches] the `id` attribute of the row; may be NULL */
	int64_t *		sort_attr;      /* [max_matches] value of the first INT sort key; may be NULL */
	mgpu_wordstat *	word_stats;     /* [n_words], per words[] entry; may be NULL */
} mgpu_result;

typedef struct mgpu_index mgpu_index;

/* ------------------------------------------------------------------------------------- */
/* index lifetime: replaces CSphIndex_VLN::Prealloc/Preread (src/sphinx.cpp:13782) for the files the
 * query path needs (.sph .spi .spd .spp .spe .spa .spm).  `device` = CUDA ordinal.
 * `rowid_base` = global rowid of local row 0 (contiguous rowid-range shards); only
 * used when packing multi-GPU merge keys. */
int				mgpu_index_open ( const char * path_prefix, int device, uint32_t rowid_base, mgpu_index ** out );
/* Lifetime: a batch keeps a pointer to its index, so free every batch of a handle before closing it. Closing a handle with live
   batches is refused (MGPU_E_BAD_QUERY, the handle stays valid); a successful close first drains the handle's own streams. A caller
   stream installed with mgpu_index_set_stream must outlive the batches prepared on it. */
int				mgpu_index_close ( mgpu_index * idx );
/* run this handle's kernels and copies on the caller's CUDA stream (cudaStream_t) instead of the handle's own one,
 * so that callers can bracket batches with their own events; NULL restores the private stream */
int				mgpu_index_set_stream ( mgpu_index * idx, void * cuda_stream );
/* Engine options of this handle, set once after open (the library never reads tuning from the environment; searchd would map
   its own config keys here, cf. CSphConfigSection in src/sphinxutils.h). Names: "plan_threads", "timing", "hot_store", "hot_div", "hot_min_uses",
   "hot_gb", "or_range_tiles", "dnf_pct", "stats", and the A/B switches "eager_hot", "force_hot", "group_neg", "or_bits", "bits_dnf", "bits_dnf_div", "or_class", "dnf_class", "and_kernel", "dnf", "chain",
   "reg_or", "jump" (0/1). Returns MGPU_E_BAD_QUERY for an unknown name or a value out of range. */
int				mgpu_index_set_option ( mgpu_index * idx, const char * name, int64_t value );
const char *	mgpu_last_error ( const mgpu_index * idx );   /* idx may be NULL: last open error */

/* index facts (CSphIndex::GetStats, schema), for callers building queries */
int64_t			mgpu_index_total_docs ( const mgpu_index * idx );
int32_t			mgpu_index_num_fields ( const mgpu_index * idx );
int32_t			mgpu_index_field_index ( const mgpu_index * idx, const char * name );   /* CSphSchema::GetFieldIndex */
const char *	mgpu_index_field_name ( const mgpu_index * idx, int32_t field );        /* CSphSchema::GetFieldName; NULL if out of range */
int32_t			mgpu_index_attr_index ( const mgpu_index * idx, const char * name );    /* CSphSchema::GetAttrIndex */
/* dictionary lookup: DiskIndexQwordSetup_c::Setup (src/sphinx.cpp:12950-13078). returns 1 if found */
int				mgpu_index_word_stats ( const mgpu_index * idx, const char * word, int64_t * docs, int64_t * hits );
/* algorithmic bytes of a word on this index: .spd extent incl. terminator + .spe extent */
int				mgpu_index_word_bytes ( const mgpu_index * idx, const char * word, int64_t * doclist_bytes, int64_t * skiplist_bytes );

/* ------------------------------------------------------------------------------------- */
/* the hot path: replaces sphCreateRanker + MatchExtended + ISphMatchSorter::Push/Flatten
 * (src/sphinxsearch.cpp:4167, src/sphinx.cpp:12190, src/sphinxsort.cpp:582-812) for a batch of
 * parsed queries against one local index.  Host buffers in, host buffers out; thread-safe on a
 * shared handle (calls are serialised per index).  Returns MGPU_OK if the batch ran; per-query
 * status in results[i].status. */
int				mgpu_search_batch ( mgpu_index * idx, const mgpu_query * queries, int n_queries, mgpu_result * results );

/* Same, but the timed region can exclude host<->device copies: prepare uploads the batch plan,
 * run launches the kernels on the resident plan (may be called repeatedly), fetch copies results
 * back.  mgpu_search_batch == prepare + run + fetch + free. */
typedef struct mgpu_batch mgpu_batch;
int				mgpu_batch_prepare ( mgpu_index * idx, const mgpu_query * queries, int n_queries, mgpu_batch ** out );
int				mgpu_batch_run ( mgpu_batch * b );            /* asynchronous on the index stream */
int				mgpu_batch_sync ( mgpu_batch * b );
int				mgpu_batch_fetch ( mgpu_batch * b, mgpu_result * results );
void			mgpu_batch_free ( mgpu_batch * b );
/* counters of the last run: kernels launched, work items, algorithmic bytes, postings */
typedef struct mgpu_batch_stats {
	int64_t			kernel_launches;
	int64_t			work_items;
	int64_t			algorithmic_bytes;   /* sum over queries of B(q) */
	int64_t			postings;            /* sum over queries of sum_t df(t) */
	int64_t			h2d_bytes;
	int64_t			d2h_bytes;
	float			eval_kernel_ms;      /* CUDA-event time of the fused eval kernel in the last run */
	float			merge_kernel_ms;
	float			hot_decode_ms;       /* K0: the batch's shared hot keywords decoded once into the dense store */
	int32_t			hot_terms;
	/* per launch class: [0] stream_kernel<512> (doc-only queries, single-level programs), [1] eval_kernel<hits> (hit-consuming
	 * queries on dense tiles), [2] and_kernel (doc-only DNF / pure AND queries), [3] stream_kernel<256> (deeper programs),
	 * [4] and_kernel<hits> (hit-consuming pure AND chains: PROXIMITY_BM25 over AND, phrase, proximity),
	 * [5] stream_kernel<512,or> (pure OR programs under BM25: bound pass + exact pass),
	 * [6] stream_kernel<512,dnf> (OR-of-AND-groups programs whose multi-keyword groups are all hot: the same passes) */
	float			class_ms[7];
	int32_t			class_queries[7];
	int64_t			class_bytes[7];      /* algorithmic bytes of the class's queries */
	/* host wall-clock of the batch: query planning, buffer setup + plan upload, result download + unpack */
	float			host_plan_ms, host_setup_ms, host_fetch_ms;
	float			host_wait_ms;        /* part of host_fetch_ms spent waiting for the kernels */
	float			host_total_ms;       /* mgpu_search_batch only: the whole call incl. freeing the batch */
	int32_t			or_kernel;           /* kernel of launch class 5: 3 = orbits_kernel (presence bitmaps), 1 = stream_kernel<512,1> */
	int64_t			hitlist_bytes;       /* .spp bytes of the matched documents' hitlists read by the hit stage */
	int64_t			attr_rows;           /* rows whose attributes the bound pass read for filters / sort keys (x attribute bytes) */
} mgpu_batch_stats;
int				mgpu_batch_get_stats ( const mgpu_batch * b, mgpu_batch_stats * out );
/* stats of the last mgpu_search_batch() call on this handle (that call frees its batch before returning) */
int				mgpu_index_last_search_stats ( const mgpu_index * idx, mgpu_batch_stats * out );

/* ------------------------------------------------------------------------------------- */
/* Rowid-range shards of one index behind one handle, one GPU per shard: replaces the distributed-local fan-out
 * RunLocalSearches (src/searchd.cpp:5596-5814: a thread per local index, each with its own sorter) + MergeAllMatches
 * (:4653-4738) + SetupLocalDF (:5869: global IDF statistics).  path_prefixes[s] is shard s = the s-th contiguous docid
 * range (its rowid base is the row count of the shards before it); devices[s] its CUDA ordinal.  One process, one host
 * thread per shard; the batch is planned once; the K keys per query and shard are exchanged with ncclSend/ncclRecv over
 * NVLink (libnccl is loaded at run time) when every shard has a GPU of its own, with stream-ordered device copies when
 * they share GPUs.  Results are those of the UNSHARDED index; mgpu_result.rowid holds GLOBAL rowids. */
typedef struct mgpu_sharded mgpu_sharded;
int				mgpu_sharded_open ( const char * const * path_prefixes, const int * devices, int n_shards, mgpu_sharded ** out );
void			mgpu_sharded_close ( mgpu_sharded * sh );
int				mgpu_sharded_search_batch ( mgpu_sharded * sh, const mgpu_query * queries, int n_queries, mgpu_result * results );
int				mgpu_sharded_set_option ( mgpu_sharded * sh, const char * name, int64_t value );   /* mgpu_index_set_option on every shard */
int64_t			mgpu_sharded_total_docs ( const mgpu_sharded * sh );
int				mgpu_sharded_word_docs ( const mgpu_sharded * sh, const char * word, int64_t * docs );   /* global df; returns 1 if found */
int				mgpu_sharded_word_stats ( const mgpu_sharded * sh, const char * word, int64_t * docs, int64_t * hits );   /* summed over the shards */
const char *	mgpu_sharded_last_error ( const mgpu_sharded * sh );   /* sh may be NULL: last open error */
typedef struct mgpu_sharded_stats {
	int32_t			n_shards;
	int32_t			nccl;                /* 1 = the last exchange went through ncclSend/ncclRecv */
	float			host_total_ms;       /* the whole mgpu_sharded_search_batch call */
	float			host_plan_ms;        /* global statistics + planning (once) */
	float			host_setup_ms;       /* shard threads: bind + upload + launch */
	float			host_wait_ms;        /* exchange, merge, download: until the results are on the host */
	float			host_fetch_ms;       /* unpacking into the caller's buffers */
	float			max_eval_kernel_ms;  /* slowest shard */
	float			max_hot_decode_ms;
	int32_t			kernel_launches;     /* all shards + the merge */
	int64_t			h2d_bytes;
	int64_t			d2h_bytes;
	int64_t			algorithmic_bytes;   /* SURVEY 8(d), summed over the shards */
	int64_t			postings;
} mgpu_sharded_stats;
int				mgpu_sharded_get_stats ( const mgpu_sharded * sh, mgpu_sharded_stats * out );

/* ------------------------------------------------------------------------------------- */
/* distributed-local merge: replaces MergeAllMatches/KillPlainDupes for disjoint rowid-range shards
 * (src/searchd.cpp:3910-3952, 4653-4738).  Each shard exports, per query, its K best packed
 * 128-bit keys {hi = sort key, lo = ~global_rowid:32 | weight:32} to DEVICE memory (for an NCCL
 * all-gather by the caller); the merge kernel selects the global K best. */
/* both calls are ASYNCHRONOUS on the index stream / the given stream (stream-ordered after mgpu_batch_run): no host sync inside */
int				mgpu_batch_export_keys ( mgpu_batch * b, void * dev_keys /* [nq][K][2] u64 */, void * dev_counts /* [nq] i32 */, void * dev_total_found /* [nq] i64 */, int K );
int				mgpu_merge_shard_keys ( int device, const void * dev_keys /* [n_shards][nq][K][2] u64 */, const void * dev_counts /* [n_shards][nq] i32 */,
					int n_shards, int nq, int K, void * dev_out_keys /* [nq][K][2] u64 */, void * dev_out_counts /* [nq] i32 */, void * stream );
/* unpack merged keys on the host side: rowid (global), weight */
void			mgpu_unpack_key ( const uint64_t key[2], uint32_t * global_rowid, int32_t * weight, uint64_t * sortkey_hi );

/* ------------------------------------------------------------------------------------- */
/* standalone doclist decode (kernel K1): decodes every posting of `word` into caller DEVICE or HOST
 * arrays; used by the parity tests of the VByte block decoder against the oracle's
 * DiskIndexQword_c::ReadNext restatement (src/sphinx.cpp:511-549). host arrays sized docs. */
int				mgpu_decode_doclist ( mgpu_index * idx, const char * word, uint32_t * rowid, uint32_t * hits, uint32_t * fields, uint64_t * hitlist_pos, int64_t capacity, int64_t * n_out );

/* ------------------------------------------------------------------------------------- */
/* Index consistency check: the checks of DiskIndexChecker_c (`indextool --check`, src/indexcheck.cpp:443-983,
 * 1292-1317) over the files the query path loads: schema, attribute row count, dead-row map size, duplicate document ids, dictionary
 * order / counts / checkpoints, every doclist decoded end to end (rowid order and bounds, hit counts, hitlist order, field masks,
 * hitlist offsets) and every skiplist recomputed.  Host only, no GPU.  Returns MGPU_OK when the check ran (also when it found
 * failures): *n_failures counts them and `report` holds the first 64 messages, one per line, in the reference's wording. */
int				mgpu_index_check ( const char * path_prefix, int64_t * n_failures, char * report, int report_len );

/* ------------------------------------------------------------------------------------- */
/* Query front-end: extended query syntax -> the flattened tree above.
 * Replaces sphParseExtendedQuery / XQParser_t::Parse (src/sphinxquery.cpp:1741-1830, 1990-2014; grammar src/sphinxquery.y) with its
 * tree fix-ups (XQParseHelper_c::FixupTree, :343-387) and the legacy match modes' rewrite (PrepareQueryEmulation,
 * src/searchd.cpp:2141-2190).  Host only, no GPU needed.  The tokenizer is the reference's default charset_table (ASCII
 * alphanumerics + '_', Cyrillic, case folded) plus CJK unigrams when ngram_cjk is set; min_word_len and stop words consume query
 * positions as in the reference (overshort_step / stopword_step).  @@relaxed, the phrase star and SENTENCE / PARAGRAPH are understood.  Not parsed: zones, exact-form '=',
 * wildcards, blended characters (MGPU_E_UNSUPPORTED where the syntax is recognised). */
enum {
	MGPU_MATCH_ALL = 0,       /* SPH_MATCH_ALL: every word; ranker SPH_RANK_PROXIMITY */
	MGPU_MATCH_ANY = 1,       /* SPH_MATCH_ANY: "words"/1; ranker SPH_RANK_MATCHANY */
	MGPU_MATCH_PHRASE = 2,    /* SPH_MATCH_PHRASE: "words"; ranker SPH_RANK_PROXIMITY */
	MGPU_MATCH_BOOLEAN = 3,   /* SPH_MATCH_BOOLEAN: extended syntax, ranker SPH_RANK_NONE */
	MGPU_MATCH_EXTENDED = 4   /* SPH_MATCH_EXTENDED / EXTENDED2 */
};
typedef struct mgpu_parser_settings {
	int32_t			n_fields;
	const char * const * field_names;   /* index schema, for @field limits (CSphSchema::GetFieldIndex) */
	int32_t			min_word_len;       /* CSphTokenizerSettings::m_iMinWordLen (0 -> 1) */
	int32_t			n_stopwords;
	const char * const * stopwords;     /* dictionary forms */
	int32_t			overshort_step;     /* CSphIndexSettings::m_iOvershortStep (reference default 1) */
	int32_t			stopword_step;      /* m_iStopwordStep (reference default 1) */
	int32_t			match_mode;         /* MGPU_MATCH_* */
	int32_t			ngram_cjk;          /* ngram_len=1 over the CJK ranges */
} mgpu_parser_settings;
typedef struct mgpu_parsed mgpu_parsed;
/* *out is set whenever settings and out are valid, also on a parse error (then it carries the message); free it with mgpu_parsed_free */
int				mgpu_parse_query ( const mgpu_parser_settings * settings, const char * text, mgpu_parsed ** out );
/* points q's tree members (nodes, children, words, root, and the ranker for the legacy match modes) at the parsed tree;
 * the pointers live until mgpu_parsed_free */
int				mgpu_parsed_fill ( const mgpu_parsed * p, mgpu_query * q );
const char *	mgpu_parsed_error ( const mgpu_parsed * p );      /* XQQuery_t::m_sParseError */
const char *	mgpu_parsed_warning ( const mgpu_parsed * p );    /* XQQuery_t::m_sParseWarning */
/* the tree as SHOW PLAN prints it (`transformed_tree`: sphExplainQuery + sph::RenderBsonPlan, src/sphinxsearch.cpp:300-335, 430-530;
 * a percent quorum shows its percentage, as there) */
const char *	mgpu_parsed_explain ( const mgpu_parsed * p );
void			mgpu_parsed_free ( mgpu_parsed * p );

/* ------------------------------------------------------------------------------------- */
/* Wire responder: the binary SphinxAPI `search` command for ONE local index, request packet in, reply packet out.
 * Replaces the server side of SEARCHD_COMMAND_SEARCH: HandleCommandSearch + ParseSearchQuery (src/searchd.cpp:6932-7000, 2201-2560),
 * the sort-mode / weight / filter setup between them and the index (src/sortsetup.cpp, src/sphinx.cpp:13903-13947) and SendResult
 * (src/searchd.cpp:3340-3510), for client protocol versions 1.29..1.33 (0x11D..0x121); also SEARCHD_COMMAND_KEYWORDS (HandleCommandKeywords:
 * the text tokenized as the index tokenizes it, with the dictionary's docs / hits on request).  The queries of one packet run as one
 * mgpu_search_batch call.  Anything the hot path has no counterpart for (group-by, expression rankers / sorts, geo anchors, select
 * lists other than "*", string and float filters, cutoff, outer order) is answered with SEARCHD_ERROR for that query, as searchd
 * answers a query it cannot run.  No sockets: the embedding daemon owns the connection (src/netreceive_api.cpp).
 * idx may be NULL: packets are parsed, described and answered with per-query errors (host-only use, tests).
 * A responder serves one packet at a time (its reply buffer is reused): one responder per connection thread; the index handle under
 * them is shared and thread-safe.
 * tokenizer may be NULL: min_word_len / overshort_step / stopword_step come from the index header, no stop words. */
typedef struct mgpu_api mgpu_api;
int				mgpu_api_create ( mgpu_index * idx, const char * path_prefix, const mgpu_parser_settings * tokenizer, mgpu_api ** out );
/* the same responder over a sharded handle (one GPU per rowid-range shard): the packet's queries run as one mgpu_sharded_search_batch,
 * attribute values come from the shard that holds the row, keyword statistics are the whole index's. path_prefixes in shard order. */
int				mgpu_api_create_sharded ( mgpu_sharded * sh, const char * const * path_prefixes, int n_shards, const mgpu_parser_settings * tokenizer, mgpu_api ** out );
/* *reply points into memory owned by the responder, valid until the next call on it or mgpu_api_free. Returns MGPU_OK whenever a
 * reply packet was produced (protocol and query errors travel inside the reply, as on the wire). */
int				mgpu_api_handle ( mgpu_api * api, const void * request, size_t request_len, const void ** reply, size_t * reply_len );
/* the queries of the last packet as SphinxQL-like text, one per line (what query_log_format=sphinxql logs for API queries) */
const char *	mgpu_api_describe_last ( const mgpu_api * api );
void			mgpu_api_free ( mgpu_api * api );

int				mgpu_abi_version ( void );

#ifdef __cplusplus
}
#endif
#endif /* MGPU_H_ */
