// Plain structs shared between the host planner (engine.cpp) and the sm_100a kernels (kernels.cu).
// HBM layout of an index (see DESIGN.md "Data layout in HBM"):
//   spd / spp / spa : the reference's own .spd/.spp/.spa bytes, uploaded verbatim (16 B zero padding after each)
//   block table     : every term's skiplist decoded once at open (+ the implicit entry 0 and the trailing
//                     partial block the reference's reader drops, src/sphinx.cpp:13056-13073), 20 B per 32 docs
#pragma once

#include <stdint.h>

namespace mgpu
{

static const int TILE_W = 2048;			///< rowids per dense tile
static const int EVAL_THREADS = 256;
static const int EVAL_WARPS = EVAL_THREADS/32;
static const int MAX_STACK = 4;			///< doc-vector stack depth of the tile program
static const int MAX_LEAVES = 16;		///< term leaves per query on the GPU path
static const int MAX_OPS = 48;
static const int MAX_FILTERS = 4;
static const int MAX_FILTER_VALUES = 16;
static const int MAX_FIELDS = 32;
static const int STAGE_BYTES = 896;		///< per-warp staging: 15 B alignment head + 32 docs * (5+5+5+10) B, rounded up to a multiple of 128
static const int MAX_PHRASE_WORDS = 16;	///< keywords per phrase/proximity node on the GPU path
static const int MAX_NWAY = 4;			///< phrase/proximity nodes per query
static const int PRE_BLOCKS = MAX_LEAVES*( TILE_W/32+2 );	///< predecode scratch per CTA: a keyword has at most TILE_W/32+2 blocks overlapping a tile
static const int MAX_GROUPS = 8;		///< AND groups of a DNF program handled by the intersection kernel
static const int NWAY_MAX_SPAN = 31;	///< max (last atom pos - first atom pos) of a phrase/proximity node: bounds the FSM state

struct DevIndex_t
{
	const uint8_t *		m_pSpd;
	const uint8_t *		m_pSpp;
	const uint32_t *	m_pSpa;
	const uint32_t *	m_pDead;		///< dead-row bitmap (.spm) or null
	const uint32_t *	m_pBlkRowid;	///< SkiplistEntry_t::m_tBaseRowIDPlus1 per block
	const uint64_t *	m_pBlkOff;		///< SkiplistEntry_t::m_iOffset
	const uint64_t *	m_pBlkHitpos;	///< SkiplistEntry_t::m_iBaseHitlistPos
	int64_t				m_iSpdLen;
	int64_t				m_iSppLen;
	uint32_t			m_uRows;
	int32_t				m_iStride;		///< .spa row stride in DWORDs
	int32_t				m_bInlineHits;
	uint32_t			m_uRowidBase;	///< global rowid of local row 0 (shards)
};

/// one query keyword occurrence = ExtTerm_T / ExtMultiAnd_T::NodeInfo_t
struct DevLeaf_t
{
	uint64_t	m_uDoclistEnd;		///< offset of the doclist's terminating zero byte
	uint32_t	m_uFirstBlk;		///< index into the block table
	uint32_t	m_nBlocks;
	uint32_t	m_nDocs;
	uint32_t	m_uQueriedFields;	///< XQLimitSpec_t field mask (fields 0..31)
	float		m_fIDF;
	uint16_t	m_uAtomPos;
	uint16_t	m_uNodePos;
	int32_t		m_iHot;				///< slot in the batch's dense hot-term store, -1 = evaluate from the compressed doclist
	int32_t		m_iTermPos;			///< ExtTermPos_T filter: low 3 bits = TermPosFilter_e (0 none, 1 field start, 2 field end, 3 both, 4 field limit), rest = m_iFieldMaxPos
	uint32_t	m_uListOff;			///< launch class 5, keywords outside the hot store: first entry of the keyword's decoded posting list (DevPostingLists_t)
	uint32_t	m_uPad;
};

/// Dense hot-term store, rebuilt by hot_decode_kernel at the start of every batch run: keywords that many queries of the
/// batch share and that sit in >= 1/8 of the rows are decoded ONCE per batch into one byte of hit count + one byte of field
/// mask per row (the GPU's take on the reference's multi-query common-subtree cache, src/searchnode.cpp:5487-5890).
struct DevHotStore_t
{
	const uint16_t *	m_pData;		///< [nHot][m_iStride]: low byte 0 = keyword absent, 1..254 = hit count, 255 = see escape list;
										///< high byte = field mask (indexes with <= 8 fields); with m_bTfClass its high nibble holds
										///< ceil ( 15*hits/(hits+1.2) ), the row's tf class for the weight bound of stream_kernel
	const uint32_t *	m_pEscape;		///< [m_nEscape][4]: hot slot, rowid, hit count, next entry of the bucket (documents with >= 255 hits)
	const int32_t *		m_pEscapeCount;	///< [1+65536]: entries, then the bucket heads (-1 = empty)
	int64_t				m_iStride;		///< rows rounded up to TILE_W
	int32_t				m_nHot;
	int32_t				m_bTfClass;		///< indexes with <= 4 fields: bits 12..15 of a row hold its tf class (consumers mask fields by the queried fields)
	const uint32_t *	m_pBits;		///< [nHot][m_nBitFields][m_iBitStride]: presence bitmap per (hot keyword, field), 1 bit per row (indexes with <= 4 fields; else null)
	int64_t				m_iBitStride;	///< words per bitmap = m_iStride/32
	int32_t				m_nBitFields;	///< fields of the index (1..4)
	int32_t				m_iPad;
	const int32_t *		m_pLvlSlot;		///< [nHot] slot in m_pLvlBits, -1 = none: keywords in >= 1/3 of the rows (the ones whose idf can be negative)
	const uint32_t *	m_pLvlBits;		///< [slots][2][m_iBitStride]: rows with >= 2 hits, rows with >= 4 hits (tf levels of orbits_kernel's penalty classes)
};

/// Decoded posting lists of the batch's non-hot keywords that launch class 5 reads (sparse_decode_kernel, once per batch run):
/// rowids ascending per keyword, value = hits (24 bits, clamped) | field mask<<24 (indexes with <= 8 fields)
struct DevPostingLists_t
{
	const uint32_t *	m_pRows;
	const uint32_t *	m_pVals;
};

struct SparseDecodeParams_t
{
	DevIndex_t				m_tIndex;
	const DevLeaf_t *		m_pTerms;		///< [nTerms] doclist descriptors; m_uListOff = where the keyword's list starts
	const uint32_t *		m_pBlkStart;	///< [nTerms+1] prefix sums of the keywords' block counts
	int32_t					m_nTerms;
	int32_t					m_iPad;
	uint32_t *				m_pRows;
	uint32_t *				m_pVals;
};

enum DevOpCode_e : uint8_t
{
	OP_TERM_SET = 0,	///< v[dst] = leaf
	OP_TERM_AND,		///< v[dst] &= leaf      (ExtAnd_c / ExtMultiAnd_T step: tfidf += , fields |=)
	OP_TERM_OR,			///< v[dst] |= leaf      (ExtOr_c)
	OP_TERM_ANDNOT,		///< v[dst] -= leaf      (ExtAndNot_c)
	OP_TERM_MAYBE,		///< v[dst] ?= leaf      (ExtMaybe_c)
	OP_VEC_AND,			///< v[dst] &= v[src]
	OP_VEC_OR,
	OP_VEC_ANDNOT,
	OP_VEC_MAYBE,
	OP_NWAY				///< v[dst] = FSM filter (phrase/proximity) over the hits of v[dst]'s leaves; arg = nway index
};

struct DevOp_t
{
	uint8_t		m_eCode;
	uint8_t		m_uDst;
	uint8_t		m_uSrc;			///< VEC ops: source level. TERM_AND: 1 + index of the first op after this AND chain (0 = none):
							///< where to resume when no candidate is left in the tile
	uint8_t		m_uAliveDst;	///< value of cnt[] meaning "alive" in v[dst] BEFORE the op
	uint8_t		m_uAliveSrc;	///< same for v[src]
	uint8_t		m_uLeaf;
	uint8_t		m_uArg;
	uint8_t		m_uAliveOut;	///< alive value of v[dst] AFTER the op
};

struct DevFilter_t
{
	int64_t		m_iMin, m_iMax;
	int64_t		m_dValues[MAX_FILTER_VALUES];
	int32_t		m_nValues;
	int32_t		m_eKind;
	int32_t		m_iDwordOff;	///< attribute locator: DWORD offset in the row
	int32_t		m_iBitCount;	///< 32 or 64
	int32_t		m_bExclude;
	int32_t		m_iPad;
};

struct DevSortKey_t
{
	int32_t		m_eKind;
	int32_t		m_iDwordOff;
	int32_t		m_iBitCount;
	int32_t		m_bDesc;
	int32_t		m_iShift;		///< position of the key's low bit inside the 64-bit packed sort key
};

/// kinds of hit-level nodes over plain keywords (the "n-way" slot of the hit stage)
enum DevNWayKind_e : int32_t
{
	NWAY_PHRASE = 0,		///< ExtNWay_T<FSMphrase_c>
	NWAY_PROXIMITY = 1,		///< ExtNWay_T<FSMproximity_c>
	NWAY_NEAR = 2,			///< ExtNWay_T<FSMmultinear_c>, src/searchnode.cpp:4080-4315 (children = keywords)
	NWAY_BEFORE = 3,		///< ExtOrder_c, src/searchnode.cpp:4657-4935 (children = keywords)
	NWAY_NOTNEAR = 4,		///< ExtNotNear_c, src/searchnode.cpp:5325-5478 (MUST keyword, NOT keyword)
	NWAY_QUORUM = 5			///< ExtQuorum_c, src/searchnode.cpp:4319-4650
};

/// phrase / proximity / NEAR / BEFORE / NOTNEAR / quorum node over plain keywords
struct DevNWay_t
{
	int32_t		m_eKind;					///< DevNWayKind_e (0 / 1 = phrase / proximity)
	int32_t		m_iOpArg;					///< "..."~N, NEAR/N, NOTNEAR/N, quorum threshold
	int32_t		m_nWords;
	int32_t		m_iQLen;					///< last atom pos - first atom pos
	int32_t		m_dLeaf[MAX_PHRASE_WORDS];	///< leaves in query (atom pos) order
	int32_t		m_dAtomPos[MAX_PHRASE_WORDS+1];
	int32_t		m_dQposDelta[NWAY_MAX_SPAN+1];	///< FSMphrase_c::m_dQposDelta (src/searchnode.cpp:3884-3899)
	uint8_t		m_dCount[MAX_PHRASE_WORDS];	///< quorum: how often the query repeats the keyword (ExtQuorum_c::TermTuple_t::m_iCount)
};

/// A device query = core + extension. The core (1.4 KB: keywords, program, weights, scalars) is what every query has; the extension
/// (1.9 KB: filters, sort keys, hit-level nodes) only exists for queries that use any of it (m_iExt >= 0) and travels in an array of its
/// own, so a 10k-query batch of plain boolean queries moves 14 MB of descriptors instead of 33 MB (host passes, upload, per-item loads).
struct DevQueryCore_t
{
	DevLeaf_t	m_dLeaves[MAX_LEAVES];
	DevOp_t		m_dOps[MAX_OPS];
	int32_t		m_dWeights[MAX_FIELDS];	///< bound field weights
	int32_t		m_nLeaves;
	int32_t		m_nOps;
	int32_t		m_nFilters;
	int32_t		m_nSortKeys;			///< 0 = relevance
	int32_t		m_nNWay;
	int32_t		m_nWeights;				///< CSphQueryContext::m_iWeights = min(fields, 32) here
	int32_t		m_eRanker;
	int32_t		m_iIndexWeight;
	int32_t		m_iMaxMatches;
	int32_t		m_uAliveRoot;			///< alive value of v[0] after the last op
	int32_t		m_bNeedHits;			///< ranker consumes hits (PROXIMITY_BM25 multi-word, WORDCOUNT) or tree has phrase/proximity
	int32_t		m_bDupes;				///< HasQwordDupes: RankerState_Proximity_fn<*,true>
	int32_t		m_iFirstItem;			///< first work item of this query
	int32_t		m_nItems;
	int32_t		m_bStateRanker;			///< ExtRanker_State_T: PROXIMITY_BM25 over >1 keyword, WORDCOUNT
	uint32_t	m_uPreMask;				///< leaves whose op scatters unconditionally (SET/OR/ANDNOT/MAYBE): decoded in the tile predecode phase
	uint32_t	m_uOrigMask;			///< leaves whose op can bring a document into the result (SET / OR operands)
	int32_t		m_bOrigHot;				///< one of those is a hot (dense) keyword: every mini-tile has to be visited
	int32_t		m_iExt;					///< index of the query's DevQueryExt_t in the batch's extension array; -1 = none (no filters, sort keys, hit-level nodes)
	int32_t		m_iMaxQpos;				///< ExtRanker_c::m_iMaxQpos (SPH04 exact-hit test)
	int32_t		m_nQwords;				///< ExtRanker_c::m_iQwords (MATCHANY phrase factor)
	int32_t		m_bPureOr;				///< the program is SET, OR, OR... over keywords (single level): eligible for the register path of stream_kernel
	int32_t		m_bWeightKey;			///< one of the sort keys is the weight (host routing)
	int32_t		m_bGroupNeg;			///< the (single) AND group ends in TERM_ANDNOT ops: `a b -c` (and_kernel rejects candidates that hold them)
	int32_t		m_iPad;
	int32_t		m_nGroups;				///< >0: the program is an OR of AND groups (DNF; 1 = pure AND): op ranges below
	uint8_t		m_dGroupOp0[MAX_GROUPS];
	uint8_t		m_dGroupOps[MAX_GROUPS];
	int32_t		m_iDriverLeaf;			///< pure AND program opened by this (sparse) keyword: tiles without its postings are skipped; -1 = none
};

struct DevQueryExt_t
{
	DevFilter_t	m_dFilters[MAX_FILTERS];
	DevSortKey_t m_dSortKeys[5];
	DevNWay_t	m_dNWay[MAX_NWAY];
};

/// the whole query as the planner builds it and as the kernels hold it in shared memory
struct DevQuery_t : DevQueryCore_t, DevQueryExt_t
{
};

#ifdef __CUDACC__
/// loads query iQuery of the batch into shared memory: the core, and the extension if the query has one (its counts are zero otherwise,
/// so nothing reads the extension's bytes). All nThreads threads of the CTA call it; the caller synchronises.
__device__ __forceinline__ void LoadQuery ( DevQuery_t & tDst, const DevQueryCore_t * pQueries, const DevQueryExt_t * pExt, uint32_t iQuery, int tid, int nThreads )
{
	const uint32_t * pSrc = reinterpret_cast<const uint32_t *>( pQueries+iQuery );
	uint32_t * pDst = reinterpret_cast<uint32_t *>( static_cast<DevQueryCore_t *>( &tDst ) );
	for ( int i=tid; i<(int)( sizeof(DevQueryCore_t)/4 ); i+=nThreads )
		pDst[i] = pSrc[i];
	const int iExt = pQueries[iQuery].m_iExt;
	if ( iExt>=0 )
	{
		const uint32_t * pSrcE = reinterpret_cast<const uint32_t *>( pExt+iExt );
		uint32_t * pDstE = reinterpret_cast<uint32_t *>( static_cast<DevQueryExt_t *>( &tDst ) );
		for ( int i=tid; i<(int)( sizeof(DevQueryExt_t)/4 ); i+=nThreads )
			pDstE[i] = pSrcE[i];
	}
}
#endif

/// one predecoded posting of the tile predecode phase
struct PreEntry_t
{
	uint32_t	m_uRowid;		///< 0xFFFFFFFF = not a posting of this tile
	float		m_fTf;			///< hits/(hits+1.2)*idf
	uint32_t	m_uFields;		///< already masked by the queried fields
	uint32_t	m_uPad;
};

/// 128-bit match key: hi = packed sort keys (bigger = better), lo = ~rowid:32 | weight:32
struct Key128_t
{
	uint64_t	m_uHi;
	uint64_t	m_uLo;
};

struct DevWorkItem_t
{
	uint32_t	m_uQuery;
	uint32_t	m_uRowLo;		///< rowid range [lo, hi) this item evaluates
	uint32_t	m_uRowHi;
	uint32_t	m_uPad;
};

struct DevItemOut_t
{
	int64_t		m_iTotalFound;
	int32_t		m_nKeys;
	int32_t		m_iPad;
};

struct EvalParams_t
{
	DevIndex_t				m_tIndex;
	const DevQueryCore_t *	m_pQueries;
	const DevQueryExt_t *	m_pQueryExt;		///< extensions of the queries with m_iExt >= 0
	const DevWorkItem_t *	m_pItems;
	int32_t					m_nItems;
	int32_t					m_iPoolCap;		///< keys per pool buffer (per CTA, two buffers)
	Key128_t *				m_pPool;		///< [gridDim.x][2][m_iPoolCap]
	Key128_t *				m_pItemKeys;	///< [m_nItems][m_iKMax]
	DevItemOut_t *			m_pItemOut;		///< [m_nItems]
	int32_t *				m_pCounter;		///< work-queue head
	int32_t					m_iKMax;		///< stride of m_pItemKeys
	int32_t					m_iPad;
	uint64_t *				m_pHitpos;		///< hit stage only: [gridDim.x][MAX_LEAVES][TILE_W] hitlist position of (leaf, tile slot)
	uint64_t *				m_pLeafTf;		///< hit stage only (eval_kernel): [gridDim.x][MAX_LEAVES][TILE_W] the keyword's own tf*idf bits | queried fields<<32
											///< on the slot: BEFORE / NOTNEAR / quorum nodes rebuild the document's TF*IDF and field mask from them
	PreEntry_t *			m_pOrList;		///< [gridDim.x][8 warps][512*MAX_LEAVES] sparse postings of the current mini-tile (register-OR path), or null
	unsigned long long *	m_pQueryThr;	///< [nQueries] shared lower bound of each query's K-th best key (hi word), zeroed per run
	PreEntry_t *			m_pPre;			///< [gridDim.x][PRE_BLOCKS*32] tile predecode scratch
	uint64_t *				m_pPreHitpos;	///< hit stage only: [gridDim.x][PRE_BLOCKS*32]
	DevHotStore_t			m_tHot;
	const int32_t *			m_pItemOrder;	///< stream_kernel: the k-th item taken from the queue is item m_pItemOrder[k] (null = k)
	unsigned long long *	m_pWork;		///< [2] work counters of the run for the roofline bookkeeping: [0] .spp bytes of the matched
											///< documents' hitlists that the hit stage read, [1] rows whose attributes the bound pass read (filters / sort keys)
	DevPostingLists_t		m_tLists;
	unsigned long long *	m_pDebug;		///< [8] work counters of the bound + exact pass kernels (option "stats"), or null
};

struct HotDecodeParams_t
{
	DevIndex_t				m_tIndex;
	const DevLeaf_t *		m_pTerms;		///< [nHot] doclist descriptors
	const uint32_t *		m_pBlkStart;	///< [nHot+1] prefix sums of the keywords' block counts (the flat block list)
	int32_t					m_nHot;
	int32_t					m_iEscapeCap;
	uint16_t *				m_pData;
	uint32_t *				m_pEscape;
	int32_t *				m_pEscapeCount;
	int64_t					m_iStride;
	int32_t					m_bTfClass;
	int32_t					m_nBitFields;	///< > 0: also build the per-field presence bitmaps
	uint32_t *				m_pBits;		///< [nHot][m_nBitFields][m_iStride/32], zeroed by the caller
	const int32_t *			m_pLvlSlot;		///< [nHot] or null
	uint32_t *				m_pLvlBits;		///< [slots][2][m_iStride/32], zeroed by the caller
};

struct MergeParams_t
{
	DevIndex_t				m_tIndex;
	const DevQueryCore_t *	m_pQueries;
	const DevQueryExt_t *	m_pQueryExt;		///< extensions of the queries with m_iExt >= 0
	int32_t					m_nQueries;
	int32_t					m_iKMax;
	const Key128_t *		m_pItemKeys;
	const DevItemOut_t *	m_pItemOut;
	Key128_t *				m_pScratch;		///< [nQueries][m_iScratchStride]
	int32_t					m_iScratchStride;
	int32_t					m_iPad;
	const int32_t *			m_pOutSlot;		///< device query -> output slot (the caller's query index)
	// outputs, [nSlots][m_iKMax]
	Key128_t *				m_pOutKeys;
	int64_t *				m_pOutDocid;
	int32_t *				m_pOutCount;	///< [nSlots]
	int64_t *				m_pOutTotal;	///< [nSlots]
};

} // namespace mgpu
