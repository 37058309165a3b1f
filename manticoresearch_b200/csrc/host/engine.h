// Host side of the GPU query engine: index loader (CSphIndex_VLN::Prealloc equivalent), query planner
// (ExtNode_i::Create + sphCreateRanker equivalent) and batch executor.  C++ in the reference's idiom;
// the only way in from outside is the C ABI in include/mgpu.h (api.cpp).
#pragma once

#include "index_format.h"
#include "../cuda/device_types.h"
#include "../../../include/mgpu.h"

#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace mgpu
{

/// dictionary entry + where the term's blocks live in the device block table
struct TermInfo_t
{
	uint32_t	m_uFirstBlk = 0;
	uint32_t	m_nBlocks = 0;
	int			m_iDocs = 0;
	int			m_iHits = 0;
	int64_t		m_iDoclistOffset = 0;
	int64_t		m_iDoclistLength = 0;	///< .spd extent incl. the terminating zero
	int64_t		m_iSkiplistBytes = 0;	///< .spe extent
	int			m_iOrdinal = 0;			///< position in the dictionary (flat per-keyword side arrays)
};

template<typename T>
struct DevBuf_T
{
	T *		m_p = nullptr;
	size_t	m_n = 0;
	cudaStream_t m_tPoolStream = nullptr;	///< set: the buffer came from the stream-ordered pool (cudaMallocAsync) of this stream
	cudaError_t Alloc ( size_t n )
	{
		Free();
		m_n = n;
		if ( !n )
			return cudaSuccess;
		return cudaMalloc ( (void**)&m_p, n*sizeof(T) );
	}
	/// per-batch buffers: stream-ordered pool allocation, no device-wide synchronisation on alloc or free
	cudaError_t AllocAsync ( size_t n, cudaStream_t tStream )
	{
		Free();
		m_n = n;
		if ( !n )
			return cudaSuccess;
		m_tPoolStream = tStream;
		return cudaMallocAsync ( (void**)&m_p, n*sizeof(T), tStream );
	}
	/// grow-only: keeps the allocation when it is already big enough (scratch reused across batches)
	cudaError_t Grow ( size_t n )
	{
		return n<=m_n ? cudaSuccess : Alloc ( n );
	}
	void Free()
	{
		if ( m_p )
		{
			if ( m_tPoolStream )
				cudaFreeAsync ( m_p, m_tPoolStream );
			else
				cudaFree ( m_p );
		}
		m_p = nullptr;
		m_n = 0;
		m_tPoolStream = nullptr;
	}
	~DevBuf_T() { Free(); }
	DevBuf_T() = default;
	DevBuf_T ( const DevBuf_T & ) = delete;
	DevBuf_T & operator= ( const DevBuf_T & ) = delete;
};

/// Engine options of one index handle (mgpu_index_set_option). The library never reads tuning from the environment: the
/// embedding process sets what it wants once after mgpu_index_open; the experiment switches exist for A/B measurements.
struct EngineOptions_t
{
	int		m_iPlanThreads = 0;			///< "plan_threads": host threads planning a batch (0 = by batch size, at most 16)
	int		m_bStats = 0;				///< "stats": print the work counters of the bound + exact pass kernels to stderr after every batch
	int		m_bTiming = 0;				///< "timing": print the host setup phases of every batch to stderr
	int		m_bHotStore = 1;			///< "hot_store": decode keywords shared by >= 2 queries once per batch into the dense store
	int		m_iHotDiv = 200;			///< "hot_div": a hot keyword sits in >= 1/hot_div of the rows
	int		m_iHotMinUses = 2;		///< "hot_min_uses": ... and used by at least this many queries of the batch
	int		m_iHotGB = 24;				///< "hot_gb": cap of the dense store
	int		m_iOrRangeTiles = 1024;		///< "or_range_tiles": rows/2048 per work item of the bound + exact pass classes
	int		m_iDnfPct = 12;				///< "dnf_pct": a group driver of the intersection kernel sits in < dnf_pct % of the rows
	int		m_bEagerHot = 1;			///< "eager_hot": one-call batches start K0 inside Prepare, on the index's second stream (0: at the head of Run)
	int		m_bGroupNeg = 2;			///< "group_neg": `a b -c` programs run on and_kernel (1) and, when all their keywords are hot, on the bitmaps (2); 0: dense tiles
	int		m_bForceHot = 1;			///< "force_hot": every keyword of a program that must run on the bitmaps (a group led by a dense keyword) enters the hot store
	int		m_bOrBits = 1;				///< "or_bits": pure OR programs run on orbits_kernel (0: stream_kernel<512,1>)
	int		m_iBitsDnfDiv = 0;			///< "bits_dnf_div": > 0 = ... only when the group's rarest keyword sits in at least 1/bits_dnf_div of the rows
	int		m_bBitsDnf = 1;				///< "bits_dnf": AND groups of hot keywords (and ORs of them) intersect their bitmaps on orbits_kernel (0: and_kernel / class 6)
	int		m_bOrClass = 1;				///< "or_class": launch class 5 exists (0: its queries take the general tile program)
	int		m_bDnfClass = 1;			///< "dnf_class": launch class 6 exists
	int		m_bAndKernel = 1;			///< "and_kernel": the intersection kernel exists
	int		m_bDnf = 1;					///< "dnf": OR-of-AND-groups programs are recognised
	int		m_bChain = 1;				///< "chain": AND chains skip ahead when a tile has no candidate left
	int		m_bRegOr = 1;				///< "reg_or": pure OR programs are recognised
	int		m_bJump = 1;				///< "jump": pure AND programs jump over tiles by their driver keyword
	bool	Set ( const char * szName, int64_t iValue );
};

class Index_c
{
public:
	std::string		m_sError;
	EngineOptions_t	m_tOpt;
	IndexHeader_t	m_tHdr;
	int				m_iDevice = 0;
	uint32_t		m_uRowidBase = 0;
	cudaStream_t	m_tStream = nullptr;		///< stream in use (own or caller's)
	cudaStream_t	m_tOwnStream = nullptr;
	cudaStream_t	m_tHotStream = nullptr;		///< second stream: a one-call batch builds its hot-term store here while the host still prepares the rest
	int				m_nSMs = 148;
	std::mutex		m_tLock;		///< serialises batches on this handle
	std::mutex		m_tCacheLock;
	std::atomic<int>	m_nLiveBatches { 0 };	///< batches prepared on this handle and not freed yet (mgpu_index_close refuses while > 0)
	std::vector<struct PlannedQuery_t> m_dPlanCache;	///< the last batch's (cleared) plan array: its memory is reused by the next batch
	mgpu_batch_stats m_tLastSearchStats {};

	std::unordered_map<std::string,TermInfo_t> m_hTerms;
	std::vector<int32_t>	m_dTermUse;		///< per keyword (by ordinal): scratch counters of Batch_c::Prepare, all zero between batches (under m_tLock)

	/// Scratch that only lives during a batch run (candidate pools, predecode lists, hit positions, the dense hot-term
	/// store): owned by the index and re-used by every batch, so a batch costs no big cudaMalloc. Contents never survive a run.
	struct RunScratch_t
	{
		DevBuf_T<Key128_t>		m_dPool;
		DevBuf_T<uint64_t>		m_dHitpos, m_dPreHitpos, m_dLeafTf;
		DevBuf_T<PreEntry_t>	m_dPre, m_dOrList;
		DevBuf_T<uint16_t>		m_dHotData;
		DevBuf_T<uint32_t>		m_dHotBits, m_dHotLvlBits;
		DevBuf_T<uint32_t>		m_dListRows, m_dListVals;	///< decoded posting lists of class 5's non-hot keywords
		DevBuf_T<uint32_t>		m_dHotEscape;
		DevBuf_T<int32_t>		m_dHotEscapeCount;
	} m_tScratch;

	/// pinned host staging for result downloads (grow-only; used under m_tLock)
	void *	m_pPinned = nullptr;
	size_t	m_nPinned = 0;
	void *	Pinned ( size_t nBytes );
	/// pinned host staging for the per-batch plan upload (grow-only; used under m_tLock, free again when Prepare returns)
	void *	m_pPinnedUp = nullptr;
	size_t	m_nPinnedUp = 0;
	void *	PinnedUpload ( size_t nBytes );

	DevBuf_T<uint8_t>	m_dSpd, m_dSpp;
	DevBuf_T<uint32_t>	m_dSpa, m_dDead;
	DevBuf_T<uint32_t>	m_dBlkRowid;
	DevBuf_T<uint64_t>	m_dBlkOff, m_dBlkHitpos;
	DevIndex_t			m_tDev {};

	int		Open ( const char * szPrefix, int iDevice, uint32_t uRowidBase );
	~Index_c();

	/// the key a query keyword has in m_hTerms: the keyword itself (dict=keywords) or its word id's key (dict=crc: the query side
	/// hashes the keyword like the indexer did, CSphDictCRC::GetWordID -> sphFNV64, src/sphinx.cpp:17318, 17569)
	std::string DictKey ( const char * szWord ) const
	{
		return m_tHdr.m_bWordDict ? std::string ( szWord ) : CrcDictKey ( WordIdFNV64 ( szWord ) );
	}
	const TermInfo_t * FindTerm ( const char * szWord ) const
	{
		auto it = m_tHdr.m_bWordDict ? m_hTerms.find ( szWord ) : m_hTerms.find ( CrcDictKey ( WordIdFNV64 ( szWord ) ) );
		return it==m_hTerms.end() ? nullptr : &it->second;
	}
	int		AttrIndex ( const char * szName ) const;
	int		FieldIndex ( const char * szName ) const;
};

/// fixed-capacity array with the few vector calls the planner uses: a plan holds no heap memory, so a batch's plan array is copied,
/// cleared and handed to another shard without 3 allocations per query
template<typename T, int N>
struct FixedVec_T
{
	T		m_d[N];
	int		m_n = 0;
	size_t	size() const				{ return (size_t)m_n; }
	T &		operator[] ( size_t i )		{ return m_d[i]; }
	const T & operator[] ( size_t i ) const { return m_d[i]; }
	const T * begin() const				{ return m_d; }
	const T * end() const				{ return m_d+m_n; }
	void	assign ( size_t n, const T & v )
	{
		m_n = (int)std::min<size_t> ( n, (size_t)N );
		for ( int i=0; i<m_n; ++i )
			m_d[i] = v;
	}
};

/// one planned query: device descriptor + host bookkeeping
struct PlannedQuery_t
{
	PlannedQuery_t() {}						///< (user-provided on purpose: a batch's plan array is not zeroed twice; the planner clears m_tDev itself)
	int				m_iStatus = MGPU_OK;
	DevQueryCore_t	m_tDev;					///< (filters, sort keys and hit-level nodes: m_tDev.m_iExt into the batch's extension array)
	int				m_nStack = 1;
	int64_t			m_iCost = 0;			///< sum of df over leaves (postings)
	int64_t			m_iAlgBytes = 0;		///< SURVEY 8(d) algorithmic bytes (doclists + skiplists)
	FixedVec_T<mgpu_wordstat,MAX_LEAVES> m_dWordStats;	///< per query keyword (queries with more than MAX_LEAVES keywords are refused)
	FixedVec_T<const TermInfo_t*,MAX_LEAVES> m_dLeafTerms;	///< dictionary entry of every leaf (null = keyword not in the index)
	FixedVec_T<int,MAX_LEAVES>	m_dLeafWord;	///< index of every leaf's keyword in mgpu_query::words
	int				m_iFirstIntKeyShift = -1;
	int				m_iFirstIntKeyBits = 0;
	bool			m_bFirstIntKeyDesc = false;
};

/// tExt: the query's filters / sort keys / hit-level nodes; returns through bHasExt whether it has any (the caller then stores tExt and sets tOut.m_tDev.m_iExt)
int		PlanQuery ( const Index_c & tIndex, const mgpu_query & tQuery, PlannedQuery_t & tOut, DevQueryExt_t & tExt, bool & bHasExt );
/// pWordIds / pTermOfId (sharded handle): the query's keywords as ids of the handle's global keyword table and this shard's dictionary
/// entry per id; without them the keywords are looked up by name
void	RebindPlan ( const Index_c & tIndex, const mgpu_query & tQuery, PlannedQuery_t & tPlan, const int32_t * pWordIds=nullptr, const TermInfo_t * const * pTermOfId=nullptr );

class Batch_c
{
public:
	Index_c *		m_pIndex = nullptr;
	cudaStream_t	m_tStream = nullptr;	///< the index's stream when the batch was prepared: everything of this batch is ordered on it
	std::string		m_sError;
	std::vector<PlannedQuery_t> m_dPlans;
	struct DevSlot_t { int m_iQuery, m_iFirstItem, m_nItems; };
	std::vector<DevSlot_t>		m_dSlots;		///< device queries = only the runnable ones, in launch order: batch query + its work items
	int							m_nDevQueries = 0;
	std::vector<int>			m_dDevToQuery;	///< device query -> batch query index
	std::vector<DevWorkItem_t>	m_dItems;
	/// launch classes: [0] doc-only queries with a single-level program (stream_kernel<512>), [1] hit-consuming queries on dense
	/// tiles (eval_kernel<true>), [2] doc-only DNF / pure AND queries led by sparse keywords (and_kernel<false>), [3] deeper doc-only
	/// programs (stream_kernel<256>), [4] hit-consuming pure AND chains incl. phrase / proximity (and_kernel<true>),
	/// [5] pure OR programs under BM25 (stream_kernel<512,1>: bound pass + exact pass), [6] OR-of-AND-groups programs whose
	/// multi-keyword groups are all hot (stream_kernel<512,2>: the same passes with gated groups)
	static const int NUM_CLASSES = 7;
	int		m_dStack[NUM_CLASSES] = { 1, 1, 1, 1, 1, 1, 1 };
	int		m_dCtas[NUM_CLASSES] = { 0, 0, 0, 0, 0, 0, 0 };
	int		m_dFirstItem[NUM_CLASSES+1] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	int		m_iKMax = 1;
	int		m_iPoolCap = 0;
	int		m_iScratchStride = 0;

	DevBuf_T<DevQueryCore_t>	m_dQ;
	std::vector<DevQueryExt_t>	m_dExt;		///< extensions of this batch's own plans (a sharded call shares the template's)
	DevBuf_T<DevQueryExt_t>	m_dExtDev;
	DevBuf_T<DevWorkItem_t>	m_dI;
	DevBuf_T<int32_t>		m_dCounter;
	std::vector<int32_t>	m_dItemOrder[2];	///< classes 5, 6: items in rowid-range-major order (index relative to the class's first item)
	DevBuf_T<int32_t>		m_dOrder[2];
	DevBuf_T<unsigned long long> m_dWork;		///< hitlist bytes read / attribute rows read (roofline bookkeeping)
	DevBuf_T<unsigned long long> m_dDebug;		///< option "stats": work counters
	DevBuf_T<unsigned long long> m_dQueryThr;	///< per device query: shared K-th-best bound of its items
	DevBuf_T<Key128_t>		m_dItemKeys, m_dScratch, m_dOutKeys;
	size_t	m_nPool = 0, m_nHitpos = 0, m_nPre = 0, m_nPreHitpos = 0;	///< what Run() needs from the index's RunScratch_t
	DevBuf_T<DevItemOut_t>	m_dItemOut;
	DevBuf_T<int64_t>		m_dOutDocid, m_dOutTotal;
	DevBuf_T<int32_t>		m_dOutCount, m_dOutSlot;

	// dense hot-term store of this batch (rebuilt by every Run(): decode once per batch instead of once per query)
	std::vector<DevLeaf_t>	m_dHotTerms;
	DevBuf_T<DevLeaf_t>		m_dHotDesc;
	std::vector<uint32_t>	m_dHotBlkStart;		///< prefix sums of the hot keywords' block counts
	DevBuf_T<uint32_t>		m_dHotBlkStartDev;
	// class 5 on orbits_kernel: its non-hot keywords are decoded once per run into plain posting lists
	std::vector<DevLeaf_t>	m_dListTerms;
	std::vector<uint32_t>	m_dListBlkStart;
	DevBuf_T<DevLeaf_t>		m_dListDesc;
	DevBuf_T<uint32_t>		m_dListBlkStartDev;
	size_t					m_nListEntries = 0;
	std::vector<int32_t>	m_dHotLvlSlot;		///< per hot keyword: slot of its tf-level bitmaps (keywords in >= 1/3 of the rows), -1 = none
	DevBuf_T<int32_t>		m_dHotLvlSlotDev;
	int						m_nHotLvl = 0;
	int						m_nHotBitFields = 0;	///< > 0: the store also holds per-field presence bitmaps (indexes with <= 4 fields)
	int						m_iOrMode = 1;			///< kernel of launch class 5: 3 = orbits_kernel, 1 = stream_kernel<512,1>
	int64_t					m_iHotStride = 0;
	int						m_iHotEscapeCap = 0;
	cudaEvent_t				m_tEvHot = nullptr, m_tEvHotDone = nullptr;
	DevHotStore_t			m_tHot {};				///< what BuildHotStore() made (pointers into the index's run scratch)
	DevPostingLists_t		m_tLists {};
	int						m_nHotLaunches = 0;
	bool					m_bHotPending = false;	///< Prepare ( bEagerHot ) has started K0 on the index's second stream: the first Run() waits for it instead of building
	cudaEvent_t				m_dEvClass[NUM_CLASSES] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	bool					m_dClassRan[NUM_CLASSES] = { false, false, false, false, false, false, false };

	cudaEvent_t		m_tEv0 = nullptr, m_tEv1 = nullptr, m_tEv2 = nullptr;
	mgpu_batch_stats m_tStats {};
	bool			m_bRan = false;

	/// pTemplate: plans made once on another shard of the same index (re-bound here instead of planning again); nMaxThreads caps the host threads
	/// bEagerHot: the caller runs the batch right away under the same lock (mgpu_search_batch, the sharded call): K0 starts inside Prepare
	/// pWordIds [ pWordOff[i] + word ] / pTermOfId: see RebindPlan
	int		Prepare ( Index_c * pIndex, const mgpu_query * pQueries, int nQueries, const std::vector<PlannedQuery_t> * pTemplate=nullptr, int nMaxThreads=0, bool bEagerHot=false,
				const int32_t * pWordIds=nullptr, const size_t * pWordOff=nullptr, const TermInfo_t * const * pTermOfId=nullptr,
				const std::vector<DevQueryExt_t> * pTemplateExt=nullptr );
	int		BuildHotStore ( cudaStream_t tStream );
	int		Run();
	int		Sync();
	int		Fetch ( mgpu_result * pResults );
	int		ExportKeys ( void * pDevKeys, void * pDevCounts, void * pDevTotal, int iK, void * pDevDocids=nullptr );
	~Batch_c();
};

// kernels.cu launchers
size_t		EvalDynSmemBytes ( int nStack );
int			EvalOccupancy ( int nStack );
cudaError_t	LaunchEval ( const EvalParams_t & P, int nStack, int nCtas, cudaStream_t tStream );	///< eval_kernel<hits>
cudaError_t	LaunchStream ( const EvalParams_t & P, int nStack, int iMode, int nCtas, cudaStream_t tStream );	///< 0 general, 1 pure OR, 2 hot DNF
int			StreamOccupancy ( int nStack, int iMode );
int			StreamOrListCap ( int iMode );
cudaError_t	LaunchAnd ( const EvalParams_t & P, bool bHits, int nCtas, cudaStream_t tStream );
int			AndOccupancy ( bool bHits );
cudaError_t	LaunchHotDecode ( const HotDecodeParams_t & P, int nCtas, cudaStream_t tStream );
cudaError_t	LaunchSparseDecode ( const SparseDecodeParams_t & P, int nCtas, cudaStream_t tStream );
cudaError_t	LaunchMerge ( const MergeParams_t & P, int nCtas, cudaStream_t tStream );
cudaError_t	LaunchShardMerge ( const Key128_t * pKeys, const int32_t * pCounts, int nShards, int nQueries, int iK,
				Key128_t * pScratch, int iScratchStride, Key128_t * pOutKeys, int32_t * pOutCounts, int nCtas, cudaStream_t tStream,
				const int64_t * pDocids=nullptr, const uint32_t * pShardBase=nullptr, int64_t * pOutDocid=nullptr );	///< + the matches' document ids
cudaError_t	LaunchDecodeDoclist ( const DevIndex_t & tIdx, const DevLeaf_t & tLeaf, uint32_t * pRowid, uint32_t * pHits, uint32_t * pFields,
				uint64_t * pHitlistPos, unsigned long long * pChecksum, int nCtas, cudaStream_t tStream );

} // namespace mgpu
