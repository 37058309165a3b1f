// K2/K3/K6/K7/K8 for doc-only queries: the warp-autonomous streaming evaluators (three instantiations of one kernel).
//
// A work item = (query, rowid range). The CTA cuts the range into 8 contiguous sub-ranges, one per warp; every warp walks
// its sub-range in 256- or 512-row mini-tiles with NO CTA-wide barrier inside a mini-tile, so that an SM has 24-32 independent
// latency chains in flight instead of one per CTA.
//
// MODE 0 (general tile program, launch classes 0 and 3): private dense vectors {tfidf, fields, cnt} in shared memory;
//   - hot keywords: 8 rows per lane straight from the batch's dense store (u16 per row);
//   - sparse keywords: a per-(warp, keyword) cursor over the resident block table; the current 32-doc block is decoded once
//     per warp into an L1/L2-resident cache and re-used by all the mini-tiles it spans; AND chains test the candidates'
//     bitmap before decoding a block (DiskIndexQword_c::HintRowID skip, src/sphinx.cpp:407-451);
//   - mini-tiles that no document-originating keyword touches are jumped over using the exact next rowid;
//   - survivors are ranked per lane and pushed into the CTA's candidate pool; warps meet at one barrier every 8 mini-tiles,
//     where the pool is compacted to K by the CTA radix select if it could overflow.
// MODE 1 / 2 (bound pass + exact pass, launch classes 5 and 6): no vectors at all. A top-K query has to COUNT every
//   matching row but only has to RANK rows that can still enter the top K: an integer pass over 8 consecutive rows per lane (one
//   128-bit load per hot keyword) yields presence, matched fields and an upper bound of the weight from the store's per-row tf
//   classes; only rows whose bound reaches the K-th best weight so far get their exact TF*IDF (one lane per row, 32 at a time).
//   MODE 1 = pure OR programs in relevance order without filters (the lean instantiation); MODE 2 adds OR-of-AND-groups programs
//   (hot groups: ANDed presence flags gate the group's sums; groups driven by one sparse keyword), filters inside the bound pass
//   and attribute / rowid sort keys. The pool is compacted on demand instead of at fixed rounds.
// Semantics per op are those of eval_kernel (ApplyTermOp): ExtTerm_T / ExtAnd_c / ExtMultiAnd_T / ExtOr_c / ExtAndNot_c /
// ExtMaybe_c of src/searchnode.cpp, TF*IDF in the reference's association order.
#pragma once
// (included from kernels.cu inside namespace mgpu: uses its DecodeBlock / ApplyTermOp / CtaSelectTopK / MakeKey helpers)

static const int OR_LIST_CAP = 512*MAX_LEAVES;	///< MODE 1 / 2: a mini-tile holds at most 512 postings of each of <=16 sparse keywords
static const int CHUNK_K = 8;				///< rows per lane handled at once (register arrays)
static const int STREAM_POOL_SLACK = 32768;	///< candidates one round (SYNC_MINIS mini-tiles per warp) can add to the pool

template<int MINI_W>
struct MiniVec_T
{
	float *		m_pTfidf;
	uint32_t *	m_pFields;
	uint8_t *	m_pCnt;
	__device__ __forceinline__ float &		Tfidf ( int v, int s )	{ return m_pTfidf[v*MINI_W+s]; }
	__device__ __forceinline__ uint32_t &	Fields ( int v, int s )	{ return m_pFields[v*MINI_W+s]; }
	__device__ __forceinline__ uint8_t &	Cnt ( int v, int s )	{ return m_pCnt[v*MINI_W+s]; }
	__device__ __forceinline__ uint32_t &	Emit ( int, int )		{ return m_pFields[0]; }	// never used (doc-only queries)
};

struct StreamShared_t
{
	DevQuery_t		m_tQ;
	SelectSmem_t	m_tSel;
	Key128_t		m_tThr;
	unsigned long long m_uTotal;
	int				m_iItem;
	int				m_iPoolCnt;
	int				m_iPoolBuf;
	uint32_t		m_dRankTab[16];
	int32_t			m_dRankUb[16];						///< bound pass: ( field-weight sum*1000 + 500 )*64 + rounding margin, per matched-field mask
	const uint16_t * m_dOpPtr[MAX_LEAVES];				///< bound pass: the op's row of the dense store (null = sparse keyword)
	uint8_t			m_dHotLeaf[MAX_LEAVES];				///< exact pass: hot keywords only (compacted, op order): the keyword's leaf
	const uint16_t * m_dHotPtr[MAX_LEAVES+4];			///< ... and its row of the dense store
	int32_t			m_nHotOps;
	uint2			m_dPosMaskUb[MAX_LEAVES];			///< bound pass, hot keywords with idf > 0: x = field-nibble mask of a row pair, y = bound per tf class
	const uint16_t * m_dPosPtr[MAX_LEAVES+4];
	uint2			m_dNegMaskUb[MAX_LEAVES];			///< bound pass, hot keywords with idf < 0: y = |idf| share per tf class, rounded down
	const uint16_t * m_dNegPtr[MAX_LEAVES];
	int32_t			m_nPosOps, m_nNegOps;
	uint8_t			m_dSparseOp[MAX_LEAVES];			///< register-OR path: ops whose keyword is walked from the compressed doclist, in op order
	int32_t			m_nSparseOps;
	uint8_t			m_dSparseUnit[MAX_LEAVES];			///< DNF mode: the AND group this sparse keyword drives (its other keywords are hot), 0xFF = a one-keyword unit
	uint8_t			m_dUnitList[MAX_LEAVES];			///< DNF mode: the unit's postings come from the mini-tile's sparse list (sparse single / sparse-driven group)
	// DNF mode: multi-keyword AND groups (hot keywords only), entries in op order; x bit 0 = negative idf
	uint2			m_dMultiMaskUb[MAX_LEAVES];
	const uint16_t * m_dMultiPtr[MAX_LEAVES];
	uint8_t			m_dMultiStart[MAX_GROUPS+1];
	int32_t			m_dMultiNeg[MAX_GROUPS];			///< 14 * sum of the group's negative shares: owed by rows the group does not match
	int32_t			m_nMultiGroups;
	const uint16_t * m_dLeafPtr[MAX_LEAVES];			///< DNF mode, exact pass: the leaf's row of the dense store (null = sparse keyword)
	int32_t			m_iNegConst;						///< 14 * sum of the negative keywords' shares: what the bound pass adds to every row on their behalf
	int32_t			m_bBound;							///< the integer weight bound is usable (no overflow, sane idf); else every present row is evaluated exactly
	float			m_dTf[256];
	uint32_t		m_dCur[EVAL_WARPS][MAX_LEAVES];		///< current block of each sparse keyword, per warp
	uint32_t		m_dCached[EVAL_WARPS][MAX_LEAVES];	///< which block sits in the warp's cache (0xFFFFFFFF = none)
	uint16_t		m_dOpStart[EVAL_WARPS][MAX_LEAVES+2];	///< register-OR path: first list entry of each op's sparse postings in this mini-tile
	uint32_t		m_dNext[EVAL_WARPS][MAX_LEAVES];	///< lower bound of the keyword's next rowid at/after the warp's position (exact once its block was examined)
	uint16_t		m_dRecStart[EVAL_WARPS][34];
	__align__(16) uint8_t m_dStage[EVAL_WARPS][STAGE_BYTES];
};

/// makes sure block b of the keyword sits decoded in the warp's cache: entries = (rowid | 0xFFFFFFFF, tf*idf, queried fields)
__device__ __forceinline__ void StreamCacheBlock ( const DevIndex_t & tIdx, const DevLeaf_t & tLeaf, uint32_t b, uint32_t * pCachedIdx,
	PreEntry_t * pCache, uint8_t * pStage, uint16_t * pRecStart, const float * pTf, int iLane )
{
	if ( *pCachedIdx==b )
		return;
	__syncwarp();
	DecodedDoc_t tDoc;
	DecodeBlock<false> ( tIdx, tLeaf, b, pStage, pRecStart, iLane, tDoc );
	const uint32_t uFields = tDoc.m_uFields & tLeaf.m_uQueriedFields;
	PreEntry_t tEntry;
	tEntry.m_uRowid = ( tDoc.m_bValid && uFields ) ? tDoc.m_uRowid : 0xFFFFFFFFu;
	// ExtTerm_T::GetDocsChunk, src/searchnode.cpp:1946
	const float fHits = __uint2float_rn ( tDoc.m_uHits );
	tEntry.m_fTf = __fmul_rn ( tDoc.m_uHits<255 ? pTf[tDoc.m_uHits & 255u] : __fdiv_rn ( fHits, __fadd_rn ( fHits, 1.2f ) ), tLeaf.m_fIDF );
	tEntry.m_uFields = uFields;
	tEntry.m_uPad = 0;
	pCache[iLane] = tEntry;
	__syncwarp();
	if ( iLane==0 )
		*pCachedIdx = b;
	__syncwarp();
}

/// moves the keyword's cursor to the first block whose last rowid is >= uRow; returns it (nBlocks = exhausted)
__device__ __forceinline__ uint32_t StreamSeek ( const uint32_t * pBase, uint32_t nBlocks, uint32_t uCur, uint32_t uRow, int iLane )
{
	// last rowid of block b = base[b+1]-1 (the skiplist entry is the previous block's last rowid + 1, src/sphinx.cpp:8447-8453)
	if ( uCur+1<nBlocks && __ldg ( pBase+uCur+1 )<=uRow )
	{
		const uint32_t u = WarpLowerBound ( pBase, uCur+1, nBlocks, uRow+1, iLane );	// first block with base > uRow
		uCur = u-1;
	}
	return uCur;
}

/// documents with >= 255 hits of a hot keyword: the real count comes from the escape list (kept out of the hot loop)
__device__ __noinline__ float HotEscapeTf ( const DevHotStore_t & tHot, int iHot, uint32_t uRowid )
{
	const float fHits = __uint2float_rn ( HotEscapeHits ( tHot, iHot, uRowid ) );
	return __fdiv_rn ( fHits, __fadd_rn ( fHits, 1.2f ) );
}

/// one hot keyword over the lane's 8 rows of the mini-tile, op code resolved at compile time (ApplyTermOp semantics)
template<int CODE>
__device__ __forceinline__ void StreamDenseOp ( float * pT, uint32_t * pF, uint8_t * pC, const uint32_t ( &dRaw )[CHUNK_K], const float * pTf,
	const DevHotStore_t & tHot, int iHot, uint32_t uRow0, uint32_t uQueried, float fIDF, uint8_t uAlive, uint8_t uAliveOut )
{
	#pragma unroll
	for ( int k=0; k<CHUNK_K; ++k )
	{
		uint32_t uHits = dRaw[k] & 255u;
		const uint32_t uFields = ( dRaw[k]>>8 ) & uQueried;
		if ( !uHits || !uFields )
		{
			if ( CODE==OP_TERM_SET )
				pC[k*32] = 0;
			continue;
		}
		float fBase = pTf[uHits];
		if ( uHits==255 )
			fBase = HotEscapeTf ( tHot, iHot, uRow0+k*32 );	// rare, out of line
		const float fTf = __fmul_rn ( fBase, fIDF );
		if ( CODE==OP_TERM_SET )
		{
			pT[k*32] = fTf; pF[k*32] = uFields; pC[k*32] = 1;
		} else if ( CODE==OP_TERM_ANDNOT )
		{
			if ( pC[k*32]==uAlive )
				pC[k*32] = 0;
		} else
		{
			const bool bAlive = pC[k*32]==uAlive;
			if ( bAlive )
			{
				pT[k*32] = __fadd_rn ( pT[k*32], fTf );
				pF[k*32] |= uFields;
				if ( CODE==OP_TERM_AND )
					pC[k*32] = uAliveOut;
			} else if ( CODE==OP_TERM_OR )
			{
				pT[k*32] = fTf; pF[k*32] = uFields; pC[k*32] = uAlive;
			}
		}
	}
}

/// MODE 0 = the general tile program. MODE 1 (ORONLY) = the class of pure OR programs under BM25 (the host routes exactly those
/// here): only the bound + exact pass path is compiled in, so the compiler does not have to share 64 registers with the general tile
/// program, relevance order, no filters. MODE 2 (DNF) = the same path with run-time options: OR-of-AND-groups programs whose
/// multi-keyword groups consist of hot keywords only, filters inside the bound pass, attribute / rowid sort keys.
template<int MINI_W, int MODE>
__global__ void __launch_bounds__ ( EVAL_THREADS, MODE ? 3 : 4 ) stream_kernel ( EvalParams_t P, int nStack )
{
	constexpr bool ORONLY = MODE!=0;
	constexpr bool DNF = MODE==2;
	constexpr int MINI_K = MINI_W/32;							///< rows per lane
	constexpr int SYNC_MINIS = STREAM_POOL_SLACK/( EVAL_WARPS*MINI_W );	///< mini-tiles per warp between two CTA barriers
	extern __shared__ __align__(16) uint8_t dDyn[];
	__shared__ StreamShared_t S;
	const int tid = threadIdx.x, iWarp = tid>>5, iLane = tid & 31;
	const DevIndex_t & tIdx = P.m_tIndex;

	MiniVec_T<MINI_W> V;
	{
		uint8_t * pMine = dDyn + (size_t)iWarp*nStack*MINI_W*9;
		V.m_pTfidf = reinterpret_cast<float *>( pMine );
		V.m_pFields = reinterpret_cast<uint32_t *>( pMine + (size_t)nStack*MINI_W*4 );
		V.m_pCnt = pMine + (size_t)nStack*MINI_W*8;
	}
	Key128_t * pPool0 = P.m_pPool + (size_t)blockIdx.x*2*P.m_iPoolCap;
	PreEntry_t * pCache0 = P.m_pPre + ( (size_t)blockIdx.x*EVAL_WARPS+iWarp )*MAX_LEAVES*32;
	uint8_t * pStage = S.m_dStage[iWarp];
	uint16_t * pRecStart = S.m_dRecStart[iWarp];
	{
		const float fHits = __uint2float_rn ( (uint32_t)tid );
		S.m_dTf[tid & 255] = __fdiv_rn ( fHits, __fadd_rn ( fHits, 1.2f ) );
	}

	while ( true )
	{
		__syncthreads();
		if ( tid==0 )
		{
			const int k = atomicAdd ( P.m_pCounter, 1 );
			S.m_iItem = ( k<P.m_nItems && P.m_pItemOrder ) ? __ldg ( P.m_pItemOrder+k ) : k;
		}
		__syncthreads();
		const int iItem = S.m_iItem;
		if ( iItem>=P.m_nItems )
			break;
		const DevWorkItem_t tItem = P.m_pItems[iItem];
		{
			LoadQuery ( S.m_tQ, P.m_pQueries, P.m_pQueryExt, tItem.m_uQuery, tid, EVAL_THREADS );
		}
		if ( tid==0 )
		{
			S.m_iPoolCnt = 0;
			S.m_iPoolBuf = 0;
			S.m_tThr.m_uHi = 0; S.m_tThr.m_uLo = 0;
			S.m_uTotal = 0;
		}
		if ( iLane<MAX_LEAVES )
		{
			S.m_dCur[iWarp][iLane] = 0;
			S.m_dCached[iWarp][iLane] = 0xFFFFFFFFu;
			S.m_dNext[iWarp][iLane] = 0;
		}
		__syncthreads();
		const DevQuery_t & q = S.m_tQ;
		const int iK = q.m_iMaxMatches;
		if ( tid<16 )
		{
			uint32_t uSum = 0;
			for ( int i=0; i<4 && i<q.m_nWeights; ++i )
				if ( tid & ( 1<<i ) )
					uSum += (uint32_t)q.m_dWeights[i];
			S.m_dRankTab[tid] = uSum;
			// Is the integer weight bound of the register-OR path usable? ( seed + field-weight sum*1000 )*index weight must stay far
			// from 2^31, every idf must be a sane number, and two rows' bound sums share a register: the sum over the keywords must
			// fit 16 bits (|idf| <= 0.5/keywords for any sane collection statistics). If not, every present row takes the exact pass.
			bool bOk = q.m_iIndexWeight>=1 && q.m_iIndexWeight<=1024 && q.m_nLeaves<=MAX_LEAVES;
			for ( int i=0; i<4 && i<q.m_nWeights; ++i )
				bOk = bOk && q.m_dWeights[i]>=0 && q.m_dWeights[i]<=250;
			float fSum = 0.0f;
			for ( int l=0; l<q.m_nLeaves && l<MAX_LEAVES; ++l )
			{
				const float fIDF = q.m_dLeaves[l].m_fIDF;
				bOk = bOk && fIDF<=1.0f && fIDF>=-1.0f;
				fSum += fabsf ( fIDF );
			}
			bOk = bOk && fSum<1.0f && P.m_tHot.m_bTfClass;
			// ( field-weight sum*1000 + 500 )*64 + a margin for the fp32 roundings; no matched field = never a candidate
			S.m_dRankUb[tid] = !tid ? -( 1<<30 ) : bOk ? (int32_t)( ( uSum*1000u + 500u )*64u + 2u ) : 0;
			if ( !tid )
				S.m_bBound = bOk ? 1 : 0;
		}
		if ( ORONLY && tid==32 )
		{
			// Per-keyword constants of the bound pass; hot keywords are also listed compactly. MODE 1: one op per keyword (pure OR).
			// MODE 2: the program is an OR of AND groups (q.m_dGroupOp0 / m_dGroupOps); one-keyword groups are handled exactly like the
			// operands of a pure OR, the keywords of bigger groups (all hot: the host's test) go to the multi-group lists.
			// The store holds c = ceil ( 15*tf ) per row (tf class). idf > 0: tf*idf <= c*idf/15, share rounded up.
			// idf < 0: tf > (c-1)/15, so tf*idf <= -(c-1)*|idf|/15, share rounded down; the bound pass adds ( 14-(c-1) )*share
			// per present row and 14*share per absent one, and the threshold moves up by the constant 14*share.
			int nHot = 0, nPos = 0, nNeg = 0, nSparse = 0, iNegConst = 0, nMulti = 0, nMultiGroups = 0;
			for ( int i=0; i<MAX_LEAVES; ++i )
			{
				S.m_dOpPtr[i] = nullptr;
				S.m_dLeafPtr[i] = nullptr;
			}
			const bool bGroups = DNF && q.m_nGroups>0;
			const int nUnits = bGroups ? q.m_nGroups : q.m_nOps;
			for ( int u=0; u<nUnits && u<MAX_LEAVES; ++u )
			{
				const int iOp0 = bGroups ? q.m_dGroupOp0[u] : u, nUnitOps = bGroups ? q.m_dGroupOps[u] : 1;
				int iGroupNeg = 0;
				// a group with ONE sparse keyword is driven by it: its postings are gathered per mini-tile, the group's other (hot)
				// keywords are looked up at those rows only and the list entry carries the whole group's exact sum (the host's test
				// admits no group with two sparse keywords)
				int nUnitSparse = 0;
				for ( int iOp=iOp0; iOp<iOp0+nUnitOps; ++iOp )
					nUnitSparse += q.m_dLeaves[q.m_dOps[iOp].m_uLeaf].m_iHot<0 ? 1 : 0;
				const bool bListUnit = nUnitSparse>0;
				S.m_dUnitList[u] = bListUnit ? 1 : 0;
				if ( nUnitOps>1 && !bListUnit )
					S.m_dMultiStart[nMultiGroups] = (uint8_t)nMulti;
				for ( int iOp=iOp0; iOp<iOp0+nUnitOps; ++iOp )
				{
					const int l = q.m_dOps[iOp].m_uLeaf;
					const DevLeaf_t & tLeaf = q.m_dLeaves[l];
					if ( tLeaf.m_iHot<0 )
					{
						S.m_dSparseUnit[nSparse] = nUnitOps>1 ? (uint8_t)u : (uint8_t)0xFF;
						S.m_dSparseOp[nSparse++] = (uint8_t)iOp;
						continue;
					}
					const float fIDF = tLeaf.m_fIDF;
					const uint16_t * pRow = P.m_tHot.m_pData + (size_t)tLeaf.m_iHot*P.m_tHot.m_iStride;
					if ( bListUnit )
					{
						S.m_dLeafPtr[l] = pRow;	// probed at the driver's postings only; its bound comes with the list entry
						continue;
					}
					const uint32_t m = tLeaf.m_uQueriedFields & 0xFu;	// this path runs for indexes with <= 4 fields only
					const uint32_t uMask = ( m<<8 ) | ( m<<24 );
					const float fShare = __fmul_rn ( fminf ( fabsf ( fIDF ), 1.0f ), 64000.0f/15.0f );
					const uint32_t uShare = fIDF>=0.0f ? (uint32_t)ceilf ( fShare )+1u : ( fShare>=2.0f ? (uint32_t)floorf ( fShare )-1u : 0u );
					if ( fIDF<0.0f )
					{
						iNegConst += 14*(int)uShare;
						iGroupNeg += 14*(int)uShare;
					}
					if ( nUnitOps>1 )
					{
						S.m_dMultiMaskUb[nMulti] = make_uint2 ( uMask | ( fIDF<0.0f ? 1u : 0u ), uShare );
						S.m_dMultiPtr[nMulti++] = pRow;
					} else if ( fIDF>=0.0f )
					{
						S.m_dPosMaskUb[nPos] = make_uint2 ( uMask, uShare );
						S.m_dPosPtr[nPos++] = pRow;
					} else
					{
						S.m_dNegMaskUb[nNeg] = make_uint2 ( uMask, uShare );
						S.m_dNegPtr[nNeg++] = pRow;
					}
					S.m_dHotPtr[nHot] = pRow;
					S.m_dHotLeaf[nHot] = (uint8_t)l;
					++nHot;
					S.m_dLeafPtr[l] = pRow;
					if ( !DNF )
						S.m_dOpPtr[iOp] = pRow;
				}
				if ( nUnitOps>1 && !bListUnit )
					S.m_dMultiNeg[nMultiGroups++] = iGroupNeg;
			}
			S.m_dMultiStart[nMultiGroups] = (uint8_t)nMulti;
			S.m_nMultiGroups = nMultiGroups;
			S.m_nHotOps = nHot;
			S.m_nSparseOps = nSparse;
			S.m_nPosOps = nPos;
			S.m_nNegOps = nNeg;
			S.m_iNegConst = iNegConst;
		}
		int iMyTotal = 0;
		uint32_t uMyAttrRows = 0;	///< rows whose attributes the filters / sort keys of the bound pass read

		// this warp's contiguous share of the item
		const uint32_t nMinis = ( tItem.m_uRowHi-tItem.m_uRowLo+MINI_W-1 )/MINI_W;
		const uint32_t uMini0 = (uint32_t)( (uint64_t)nMinis*iWarp/EVAL_WARPS ), uMini1 = (uint32_t)( (uint64_t)nMinis*( iWarp+1 )/EVAL_WARPS );
		// rounds of SYNC_MINIS mini-tiles per warp between barriers, after two warm-up rounds of ONE mini-tile each: the CTA learns a
		// K-th-best threshold after 4K rows instead of 32K
		const uint32_t nPerWarp = ( nMinis+EVAL_WARPS-1 )/EVAL_WARPS;
		const uint32_t nRounds = 2 + ( ( nPerWarp>2 ? nPerWarp-2 : 0 ) + SYNC_MINIS-1 )/SYNC_MINIS + 1;	// +1: shares differ by one mini-tile
		uint32_t uMini = uMini0;
		const uint8_t uAliveRoot = (uint8_t)q.m_uAliveRoot;
		const bool bFastRank = q.m_eRanker==1 && !q.m_nFilters && !q.m_nSortKeys && q.m_nWeights<=4 && !tIdx.m_pDead;
		const bool bRegOr = ORONLY;	// (the host's test: Batch_c::Prepare; bFastRank && pure OR)
		const bool bAnyEscape = P.m_tHot.m_pEscapeCount && __ldg ( P.m_tHot.m_pEscapeCount )!=0;	// documents with >= 255 hits of a hot keyword exist

		__syncthreads();	// S.m_nHotOps and friends
		// register-OR path state: the warp's queue of candidate rows waiting for an exact pass, the next row a sparse keyword can touch
		uint32_t * pQueue = reinterpret_cast<uint32_t *>( dDyn + (size_t)iWarp*nStack*MINI_W*9 ) + MINI_W + 64;	// [32+256], after overlay + pCand
		int nQueue = 0;
		uint32_t uNextSparse = 0;
		const int nOps = q.m_nOps, nHotOps = S.m_nHotOps, nPosOps = S.m_nPosOps, nNegOps = S.m_nNegOps, nSparseOps = S.m_nSparseOps, nMultiGroups = S.m_nMultiGroups;
		const int iIndexWeight = q.m_iIndexWeight;
		// filters and / or attribute sort keys (MODE 2 only): rows are filtered in the bound pass, keys compared whole
		const bool bAttr = DNF && ( q.m_nFilters || q.m_nSortKeys );
		const bool bGroups = DNF && q.m_nGroups>0;	// units of the OR fold = AND groups (else: every op is a one-keyword unit)

		// ranks one evaluated row and pushes it if it beats the K-th best key so far (one row per lane)
		auto fnRankPush = [&] ( bool bRow, float fT, uint32_t uF, uint32_t uRow, Key128_t * pPool, const Key128_t & tThr )
		{
			// seed weight src/sphinxsearch.cpp:1070, ExtRanker_WeightSum_c :1112-1129
			const int iSeed = __float2int_rz ( __fmul_rn ( __fadd_rn ( fT, 0.5f ), 1000.0f ) );
			const uint32_t uRank = uF ? S.m_dRankTab[uF & 15u] : 1u;
			const uint32_t uW = ( (uint32_t)iSeed + uRank*1000u )*(uint32_t)iIndexWeight;
			Key128_t tKey;
			if ( bAttr )
				tKey = MakeKey ( tIdx, q, bRow ? uRow : 0u, (int)uW );	// (filters were applied before the row was queued)
			else
			{
				tKey.m_uHi = (uint64_t)( uW ^ 0x80000000u )<<32;
				tKey.m_uLo = ( (uint64_t)( ~( uRow+tIdx.m_uRowidBase ) )<<32 ) | uW;
			}
			const bool bPush = bRow && !KeyLess ( tKey, tThr );
			const unsigned m = __ballot_sync ( FULL_MASK, bPush );
			if ( m )
			{
				int iSlot = 0;
				if ( iLane==0 )
					iSlot = atomicAdd ( &S.m_iPoolCnt, __popc ( m ) );
				iSlot = __shfl_sync ( FULL_MASK, iSlot, 0 );
				if ( bPush )
					pPool[iSlot + __popc ( m & ( ( 1u<<iLane )-1u ) )] = tKey;
			}
		};
		// exact TF*IDF of one queued row per lane from the hot keywords alone (the row holds no sparse posting), in op order
		auto fnExactHot = [&] ( bool bAct, uint32_t uRow, Key128_t * pPool, const Key128_t & tThr )
		{
			float fT = 0.0f;
			uint32_t uF = 0;
			bool bPres = false;
			for ( int h0=0; h0<nHotOps; h0+=4 )
			{
				uint32_t dRaw[4];
				#pragma unroll
				for ( int i=0; i<4; ++i )
					dRaw[i] = ( bAct && h0+i<nHotOps ) ? __ldg ( S.m_dHotPtr[h0+i]+uRow ) : 0u;
				#pragma unroll
				for ( int i=0; i<4; ++i )
				{
					const uint32_t uHits = dRaw[i] & 255u;
					if ( !uHits )
						continue;
					const DevLeaf_t & tLeaf = q.m_dLeaves[S.m_dHotLeaf[h0+i]];
					const uint32_t uFields = ( dRaw[i]>>8 ) & tLeaf.m_uQueriedFields;
					if ( !uFields )
						continue;
					float fBase = S.m_dTf[uHits];
					if ( bAnyEscape && uHits==255 )
						fBase = HotEscapeTf ( P.m_tHot, tLeaf.m_iHot, uRow );
					const float fTf = __fmul_rn ( fBase, tLeaf.m_fIDF );
					// ExtOr_c: both sides -> sum, one side -> copy (src/searchnode.cpp:3486-3504)
					fT = bPres ? __fadd_rn ( fT, fTf ) : fTf;
					uF |= uFields;
					bPres = true;
				}
			}
			fnRankPush ( bAct && bPres, fT, uF, uRow, pPool, tThr );
		};

		// DNF mode: exact TF*IDF of one candidate row per lane, unit by unit in program order. A group counts when all its keywords sit on
		// the row: its sum is folded keyword by keyword (ExtAnd_c / ExtMultiAnd_T: left + right), the groups' sums are folded in group
		// order (ExtOr_c, src/searchnode.cpp:3486-3504). sSlot = the row's slot in the current mini-tile when its sparse postings are at
		// hand in pList (-1: the row holds none).
		auto fnExactDnf = [&] ( bool bAct, uint32_t uRow, int sSlot, const PreEntry_t * pList, Key128_t * pPool, const Key128_t & tThr )
		{
			float fT = 0.0f;
			uint32_t uF = 0;
			bool bPres = false;
			int iSp = 0;
			const int nUnits = bGroups ? q.m_nGroups : nOps;
			for ( int g=0; g<nUnits; ++g )
			{
				const int iOp0 = bGroups ? q.m_dGroupOp0[g] : g, nGroupOps = bGroups ? q.m_dGroupOps[g] : 1;
				float fG = 0.0f;
				uint32_t uFG = 0;
				bool bAll = true, bFirst = true;
				if ( S.m_dUnitList[g] )
				{
					// sparse single or sparse-driven group: the mini-tile's list holds the unit's exact sum at the rows it matches
					bAll = false;
					if ( sSlot>=0 )
					{
						const int iTo = S.m_dOpStart[iWarp][iSp+1];
						for ( int e=S.m_dOpStart[iWarp][iSp]; e<iTo; ++e )
						{
							const PreEntry_t tEntry = pList[e];	// same address in every lane: a broadcast
							if ( bAct && (int)tEntry.m_uRowid==sSlot )
							{
								fG = tEntry.m_fTf;
								uFG = tEntry.m_uFields;
								bAll = true;
								bFirst = false;
							}
						}
					}
					++iSp;
				} else
				for ( int iOp=iOp0; iOp<iOp0+nGroupOps; ++iOp )
				{
					const DevLeaf_t & tLeaf = q.m_dLeaves[q.m_dOps[iOp].m_uLeaf];
					const uint16_t * pRow = S.m_dLeafPtr[q.m_dOps[iOp].m_uLeaf];
					const uint32_t uRaw = ( bAct && pRow ) ? __ldg ( pRow+uRow ) : 0u;
					const uint32_t uHits = uRaw & 255u;
					const uint32_t uFields = ( uRaw>>8 ) & tLeaf.m_uQueriedFields;
					if ( !uHits || !uFields )
					{
						bAll = false;
						continue;
					}
					float fBase = S.m_dTf[uHits];
					if ( bAnyEscape && uHits==255 )
						fBase = HotEscapeTf ( P.m_tHot, tLeaf.m_iHot, uRow );
					const float fTf = __fmul_rn ( fBase, tLeaf.m_fIDF );
					fG = bFirst ? fTf : __fadd_rn ( fG, fTf );
					uFG |= uFields;
					bFirst = false;
				}
				if ( bAll && !bFirst )
				{
					fT = bPres ? __fadd_rn ( fT, fG ) : fG;
					uF |= uFG;
					bPres = true;
				}
			}
			fnRankPush ( bAct && bPres, fT, uF, uRow, pPool, tThr );
		};

		// Register-OR class: no fixed rounds. A fresh K-th-best bound is what keeps rows out of the exact pass, so the pool is compacted
		// whenever it holds iTrigger keys: every warp looks at the level before each mini-tile and comes to the barrier when it is reached
		// (a warp adds at most 512+288 keys in between, far below the pool's slack); after the warm-up pushes are rare and so are barriers.
		const int iTrigger = min ( 2*iK+1024, iK+16384 );
		for ( uint32_t uRound=0; ORONLY || uRound<nRounds; ++uRound )
		{
			// all warps meet here; compact the candidate pool if this round could overflow it
			__syncthreads();
			const int iPoolNow = S.m_iPoolCnt;
			// nobody pushes before everybody has read the level
			if constexpr ( ORONLY )
			{
				if ( !__syncthreads_or ( uMini<uMini1 ) )
					break;	// every warp has finished its share
			} else
				__syncthreads();
			if ( ORONLY ? iPoolNow>=iTrigger : ( iPoolNow+STREAM_POOL_SLACK>P.m_iPoolCap || ( uRound<=2 && iPoolNow>iK ) ) )
			{
				Key128_t * pIn = pPool0 + (size_t)S.m_iPoolBuf*P.m_iPoolCap;
				Key128_t * pOut = pPool0 + (size_t)( S.m_iPoolBuf^1 )*P.m_iPoolCap;
				Key128_t tNewThr = CtaSelectTopK ( pIn, iPoolNow, iK, pOut, S.m_tSel );
				if ( tid==0 )
				{
					S.m_tThr = tNewThr;
					S.m_iPoolCnt = iK;
					S.m_iPoolBuf ^= 1;
					// K keys of this item are >= tNewThr, so the query's global K-th best is too: share the bound with the other items
					atomicMax ( P.m_pQueryThr+tItem.m_uQuery, (unsigned long long)tNewThr.m_uHi );
				}
				__syncthreads();
			}
			Key128_t tThr = S.m_tThr;
			{
				// another item of this query may already know a better lower bound of the K-th best key (hi word; lo = 0 keeps it a bound)
				const unsigned long long uShared = *( (volatile unsigned long long *)( P.m_pQueryThr+tItem.m_uQuery ) );
				if ( uShared>tThr.m_uHi )
				{
					tThr.m_uHi = uShared;
					tThr.m_uLo = 0;
				}
			}
			Key128_t * pPool = pPool0 + (size_t)S.m_iPoolBuf*P.m_iPoolCap;
			uint32_t uThrWx, uThrRow;
			int iThrFx, iThrFxTie;
			Key128_t tThrCur;	// the warp's current bound of the K-th best key (attribute-sorted queries compare whole keys)
			auto fnSetThr = [&] ( uint64_t uHiWord, uint64_t uLoWord )
			{
				tThrCur.m_uHi = uHiWord;
				tThrCur.m_uLo = uLoWord;
				uThrWx = (uint32_t)( uHiWord>>32 );
				uThrRow = ~(uint32_t)( uLoWord>>32 );
				// threshold of the bound pass in its own fixed point: ( bound>>6 )*index weight >= K-th best weight
				iThrFx = iThrFxTie = -( 1<<29 );	// no threshold yet / no usable bound: every present row is a candidate
				if ( bRegOr && S.m_bBound && uThrWx && !q.m_nSortKeys )	// (attribute sort: the weight is no part of the key)
				{
					const int iW = q.m_iIndexWeight;
					const long long iThr = (long long)(int)( uThrWx ^ 0x80000000u );
					long long iQ = iThr>=0 ? ( iThr+iW-1 )/iW : -( ( -iThr )/iW );	// ceil ( thr / index weight )
					iQ = iQ<-( 1<<22 ) ? -( 1<<22 ) : iQ>( 1<<24 ) ? ( 1<<24 ) : iQ;
					iThrFx = (int)( iQ*64 ) + S.m_iNegConst;
					iThrFxTie = iW==1 ? iThrFx+64 : iThrFx;	// a tie with the K-th best weight only counts at a lower rowid
				}
			};
			fnSetThr ( tThr.m_uHi, tThr.m_uLo );
			const uint32_t uRoundEnd = ORONLY ? uMini1 : min ( uMini1, uMini0 + ( uRound<2 ? uRound+1 : 2+( uRound-1 )*SYNC_MINIS ) );

			while ( uMini<uRoundEnd )
			{
				if constexpr ( ORONLY )
				{
					if ( *( (volatile int *)&S.m_iPoolCnt )>=iTrigger )
						break;	// time to compact the pool: all warps meet at the barrier above
					if ( ( uMini & 7u )==7u )
					{
						// another item of this query may have raised the shared lower bound of the K-th best key meanwhile
						const unsigned long long uShared = *( (volatile unsigned long long *)( P.m_pQueryThr+tItem.m_uQuery ) );
						if ( uShared>tThrCur.m_uHi )
							fnSetThr ( uShared, 0 );
					}
				}
				const uint32_t uLo = tItem.m_uRowLo + uMini*MINI_W;
				const uint32_t uHi = min ( uLo+(uint32_t)MINI_W, tItem.m_uRowHi );

				// Does any document-originating keyword (SET / OR operand) touch this mini-tile? If none is hot, find the exact
				// next rowid any of them holds and jump there.
				if ( !q.m_bOrigHot )
				{
					uint32_t uNext = 0xFFFFFFFFu;
					for ( uint32_t m=q.m_uOrigMask; m; m&=m-1 )
						uNext = min ( uNext, S.m_dNext[iWarp][__ffs ( m )-1] );
					if ( uNext>=uHi )
					{
						// nothing can match before uNext
						if ( uNext==0xFFFFFFFFu || uNext>=tItem.m_uRowHi )
							uMini = uMini1;
						else
							uMini = max ( uMini+1, ( uNext-tItem.m_uRowLo )/MINI_W );
						continue;
					}
				}

				// Pure OR programs under BM25 relevance (the bulk of class 0): the whole tile program runs in REGISTERS, 8 rows per lane at
				// a time. Sparse keywords' postings of this mini-tile are first gathered into a short per-warp list (in op order); hot keywords
				// are read straight from the dense store; nothing touches the shared-memory vectors.
				if constexpr ( ORONLY )
				{
					// Sparse keywords' postings of this mini-tile are first gathered into a short per-warp list (in op order). Then the
					// mini-tile is evaluated in two passes of 256 rows, 8 CONSECUTIVE rows per lane (one 128-bit load per hot keyword):
					//  A. bound pass, integers only: OR of the matched-field masks, exact presence count (total_found), and an upper bound of
					//     the weight: hot keywords contribute ( tf class of the row )*( idf share ), sparse postings their exact tf*idf rounded up;
					//  B. exact pass, only for rows whose bound reaches the K-th best weight so far: TF*IDF in the reference's order, one
					//     lane per candidate row. Rows holding a sparse posting are evaluated right away (they need the mini-tile's list);
					//     the others wait in a per-warp queue until 32 of them fill a pass.
					// Rows at/after the item's end are never present: the dense store is zero there and the sparse list holds [uLo,uHi) only.
					PreEntry_t * pList = P.m_pOrList + ( (size_t)blockIdx.x*EVAL_WARPS+iWarp )*OR_LIST_CAP;
					int nList = 0;
					if ( uNextSparse<uHi )
					{
						for ( int iSp=0; iSp<nSparseOps; ++iSp )
						{
							const int l = q.m_dOps[S.m_dSparseOp[iSp]].m_uLeaf;
							const DevLeaf_t & tLeaf = q.m_dLeaves[l];
							if ( iLane==0 )
								S.m_dOpStart[iWarp][iSp] = (uint16_t)nList;
							if ( S.m_dNext[iWarp][l]>=uHi )
								continue;
							const uint32_t * pBase = tIdx.m_pBlkRowid + tLeaf.m_uFirstBlk;
							uint32_t b = tLeaf.m_nBlocks ? StreamSeek ( pBase, tLeaf.m_nBlocks, S.m_dCur[iWarp][l], uLo, iLane ) : 0;
							uint32_t uNextRow = 0xFFFFFFFFu;
							while ( b<tLeaf.m_nBlocks )
							{
								const uint32_t uBase = __ldg ( pBase+b );
								if ( uBase>=uHi )
								{
									uNextRow = uBase;
									break;
								}
								const uint32_t uNextBase = b+1<tLeaf.m_nBlocks ? __ldg ( pBase+b+1 ) : 0xFFFFFFFFu;
								StreamCacheBlock ( tIdx, tLeaf, b, &S.m_dCached[iWarp][l], pCache0+l*32, pStage, pRecStart, S.m_dTf, iLane );
								PreEntry_t tEntry = pCache0[l*32+iLane];
								bool bIn = tEntry.m_uRowid>=uLo && tEntry.m_uRowid<uHi;
								if constexpr ( DNF )
								{
									const int iUnit = S.m_dSparseUnit[iSp];
									if ( iUnit!=0xFF )
									{
										// this keyword drives an AND group: look its hot keywords up at the posting's row and fold the group's
										// sum in op order (ExtAnd_c: left + right); the entry then stands for the whole group
										const int iOp0 = q.m_dGroupOp0[iUnit], nGroupOps = q.m_dGroupOps[iUnit];
										float fG = 0.0f;
										uint32_t uFG = 0;
										bool bFirst = true;
										for ( int iOp=iOp0; iOp<iOp0+nGroupOps; ++iOp )
										{
											const int iLeaf = q.m_dOps[iOp].m_uLeaf;
											float fTf = tEntry.m_fTf;
											uint32_t uFl = tEntry.m_uFields;
											if ( iLeaf!=l )
											{
												const DevLeaf_t & tOther = q.m_dLeaves[iLeaf];
												const uint32_t uRaw = bIn ? __ldg ( S.m_dLeafPtr[iLeaf]+tEntry.m_uRowid ) : 0u;
												const uint32_t uHits = uRaw & 255u;
												uFl = ( uRaw>>8 ) & tOther.m_uQueriedFields;
												if ( !uHits || !uFl )
												{
													bIn = false;
													continue;
												}
												float fBase = S.m_dTf[uHits];
												if ( bAnyEscape && uHits==255 )
													fBase = HotEscapeTf ( P.m_tHot, tOther.m_iHot, tEntry.m_uRowid );
												fTf = __fmul_rn ( fBase, tOther.m_fIDF );
											}
											fG = bFirst ? fTf : __fadd_rn ( fG, fTf );
											uFG |= uFl;
											bFirst = false;
										}
										tEntry.m_fTf = fG;
										tEntry.m_uFields = uFG;
									}
								}
								const unsigned m = __ballot_sync ( FULL_MASK, bIn );
								if ( bIn )
								{
									tEntry.m_uRowid -= uLo;	// slot inside the mini-tile
									// the posting's share of the weight bound: its exact tf*idf, 6 fractional bits, rounded up
									tEntry.m_uPad = tEntry.m_fTf>0.0f ? (uint32_t)ceilf ( __fmul_rn ( fminf ( tEntry.m_fTf, 1.0f ), 64000.0f ) )+1u : 0u;
									pList[nList + __popc ( m & ( ( 1u<<iLane )-1u ) )] = tEntry;
								}
								nList += __popc ( m );
								if ( uNextBase>uHi )
								{
									const uint32_t r = pCache0[l*32+iLane].m_uRowid;
									uint32_t uMin = ( r!=0xFFFFFFFFu && r>=uHi ) ? r : 0xFFFFFFFFu;
									#pragma unroll
									for ( int iStep=16; iStep; iStep>>=1 )
										uMin = min ( uMin, __shfl_xor_sync ( FULL_MASK, uMin, iStep ) );
									uNextRow = min ( uMin, uNextBase );
									break;
								}
								++b;
							}
							__syncwarp();
							if ( iLane==0 )
							{
								S.m_dCur[iWarp][l] = b;
								S.m_dNext[iWarp][l] = uNextRow;
							}
						}
						if ( iLane==0 )
							S.m_dOpStart[iWarp][nSparseOps] = (uint16_t)nList;
						__syncwarp();
						// the next mini-tile any sparse keyword can touch
						uNextSparse = 0xFFFFFFFFu;
						for ( int iSp=0; iSp<nSparseOps; ++iSp )
							uNextSparse = min ( uNextSparse, S.m_dNext[iWarp][q.m_dOps[S.m_dSparseOp[iSp]].m_uLeaf] );
					}

					// the next mini-tile's rows of the hot keywords: 8 lines of 128 B each, asked for one mini-tile ahead
					if ( uLo+2*MINI_W<=tItem.m_uRowHi )
						for ( int h=iLane>>3; h<nHotOps; h+=4 )
							asm volatile ( "prefetch.global.L2 [%0];" :: "l" ( S.m_dHotPtr[h]+uLo+MINI_W+( iLane & 7 )*64 ) );

					uint32_t * pOv = reinterpret_cast<uint32_t *>( dDyn + (size_t)iWarp*nStack*MINI_W*9 );	// [512] sparse overlay: bound<<8 | fields
					uint8_t * pCand = reinterpret_cast<uint8_t *>( pOv + MINI_W );							// [256] candidate rows of the chunk
					if ( nList )
					{
						const uint4 tZero = make_uint4 ( 0, 0, 0, 0 );
						#pragma unroll
						for ( int i=0; i<MINI_W/128; ++i )
							reinterpret_cast<uint4 *>( pOv )[i*32+iLane] = tZero;
						__syncwarp();
						for ( int e=iLane; e<nList; e+=32 )
						{
							const PreEntry_t tEntry = pList[e];
							atomicAdd ( pOv+tEntry.m_uRowid, tEntry.m_uPad<<8 );
							atomicOr ( pOv+tEntry.m_uRowid, tEntry.m_uFields & 0xFFu );
						}
						__syncwarp();
					}

					#pragma unroll 1
					for ( int c=0; c<MINI_W/256; ++c )
					{
						const uint32_t uRowC = uLo + c*256 + iLane*8;	// the first of this lane's 8 rows
						uint32_t dF[4] = { 0, 0, 0, 0 };					// two rows per word: field masks at bits 8..11 and 24..27
						uint64_t dUb[4] = { 0, 0, 0, 0 };					// two rows per word: 16-bit bound sums at bits 12..27 and 28..43
						// four keywords at a time, all four 128-bit loads in flight before the first use
						for ( int h0=0; h0<nPosOps; h0+=4 )
						{
							uint4 dRaw[4];
							#pragma unroll
							for ( int i=0; i<4; ++i )
								if ( h0+i<nPosOps )
									dRaw[i] = __ldg ( reinterpret_cast<const uint4 *>( S.m_dPosPtr[h0+i]+uRowC ) );
							#pragma unroll
							for ( int i=0; i<4; ++i )
								if ( h0+i<nPosOps )
								{
									const uint2 tMaskUb = S.m_dPosMaskUb[h0+i];
									const uint32_t dW[4] = { dRaw[i].x, dRaw[i].y, dRaw[i].z, dRaw[i].w };
									#pragma unroll
									for ( int j=0; j<4; ++j )
									{
										dF[j] |= dW[j] & tMaskUb.x;	// a present row has a non-empty field mask
										dUb[j] += (uint64_t)( dW[j] & 0xF000F000u )*tMaskUb.y;	// tf classes at bits 12..15 / 28..31: one IMAD.WIDE per row pair
									}
								}
						}
						#pragma unroll 2
						for ( int h=0; h<nNegOps; ++h )
						{
							const uint4 tRaw = __ldg ( reinterpret_cast<const uint4 *>( S.m_dNegPtr[h]+uRowC ) );
							const uint2 tMaskUb = S.m_dNegMaskUb[h];
							const uint32_t dW[4] = { tRaw.x, tRaw.y, tRaw.z, tRaw.w };
							#pragma unroll
							for ( int j=0; j<4; ++j )
							{
								const uint32_t x = dW[j] & tMaskUb.x;
								dF[j] |= x;
								// nibble+15 carries into bit 12 / 28 iff the row matches a queried field: 14 - ( class-1 ) if so, else 14
								const uint32_t uPres = ( x+0x0F000F00u ) & 0x10001000u;
								const uint32_t uClass = uPres*15u & dW[j];	// class bits of the present rows only
								dUb[j] += (uint64_t)( 0xE000E000u + uPres - uClass )*tMaskUb.y;
							}
						}
						if constexpr ( DNF )
						{
							// multi-keyword AND groups: presence flags are ANDed over the group's keywords, the group's field masks and bound
							// sums only count where the whole group sits on the row
							#pragma unroll 1
							for ( int g=0; g<nMultiGroups; ++g )
							{
								uint32_t dA[4] = { 0x10001000u, 0x10001000u, 0x10001000u, 0x10001000u }, dFg[4] = { 0, 0, 0, 0 };
								uint64_t dUg[4] = { 0, 0, 0, 0 };
								const int iTo = S.m_dMultiStart[g+1];
								#pragma unroll 2
								for ( int h=S.m_dMultiStart[g]; h<iTo; ++h )
								{
									const uint4 tRaw = __ldg ( reinterpret_cast<const uint4 *>( S.m_dMultiPtr[h]+uRowC ) );
									const uint2 tMaskUb = S.m_dMultiMaskUb[h];
									const uint32_t uMask = tMaskUb.x & ~1u;
									const bool bNeg = ( tMaskUb.x & 1u )!=0;
									const uint32_t dW[4] = { tRaw.x, tRaw.y, tRaw.z, tRaw.w };
									#pragma unroll
									for ( int j=0; j<4; ++j )
									{
										const uint32_t x = dW[j] & uMask;
										const uint32_t z = ( x+0x0F000F00u ) & 0x10001000u;	// bit 12 / 28: the keyword sits on the row
										dA[j] &= z;
										dFg[j] |= x;
										const uint32_t v = bNeg ? ( 0xE000E000u + z - ( z*15u & dW[j] ) ) : ( dW[j] & 0xF000F000u );
										dUg[j] += (uint64_t)v*tMaskUb.y;
									}
								}
								const uint32_t uOwed = (uint32_t)S.m_dMultiNeg[g];
								#pragma unroll
								for ( int j=0; j<4; ++j )
								{
									const uint32_t a = dA[j];
									dF[j] |= dFg[j] & ( ( a>>4 )*15u );
									const uint64_t uGate = (uint64_t)( ( a & 0x1000u )*0xFFFFu ) | ( (uint64_t)( ( a>>28 )*0xFFFFu )<<28 );
									dUb[j] += ( dUg[j] & uGate ) + (uint64_t)( a ^ 0x10001000u )*uOwed;
								}
							}
						}
						uint32_t uSparseRows = 0;
						if ( nList )
						{
							const uint4 tOv0 = reinterpret_cast<const uint4 *>( pOv )[c*64+iLane*2], tOv1 = reinterpret_cast<const uint4 *>( pOv )[c*64+iLane*2+1];
							const uint32_t dO[CHUNK_K] = { tOv0.x, tOv0.y, tOv0.z, tOv0.w, tOv1.x, tOv1.y, tOv1.z, tOv1.w };
							if ( tOv0.x | tOv0.y | tOv0.z | tOv0.w | tOv1.x | tOv1.y | tOv1.z | tOv1.w )
							{
								#pragma unroll
								for ( int j=0; j<4; ++j )
								{
									dUb[j] += ( (uint64_t)( dO[2*j]>>8 )<<12 ) + ( (uint64_t)( dO[2*j+1]>>8 )<<28 );
									dF[j] |= ( ( dO[2*j] & 0xFu )<<8 ) | ( ( dO[2*j+1] & 0xFu )<<24 );
								}
								#pragma unroll
								for ( int k=0; k<CHUNK_K; ++k )
									if ( dO[k] )
										uSparseRows |= 1u<<k;
							}
						}

						const int iThrLane = uRowC+tIdx.m_uRowidBase<=uThrRow ? iThrFx : iThrFxTie;	// (conservative: this lane's first row)
						uint32_t uCand = 0, uPresent = 0;
						#pragma unroll
						for ( int j=0; j<4; ++j )
						{
							const int iLo = (int)( (uint32_t)( dUb[j]>>12 ) & 0xFFFFu ) + S.m_dRankUb[( dF[j]>>8 ) & 15u];
							const int iHi = (int)( (uint32_t)( dUb[j]>>28 ) & 0xFFFFu ) + S.m_dRankUb[( dF[j]>>24 ) & 15u];
							if ( iLo>=iThrLane )
								uCand |= 1u<<( 2*j );
							if ( iHi>=iThrLane )
								uCand |= 2u<<( 2*j );
							uPresent |= ( ( ( dF[j] & 0x0F000F00u )+0x0F000F00u ) & 0x10001000u )>>j;
						}
						if ( bAttr )
						{
							// filters (EarlyReject, src/sphinx.cpp:11903) run on every present row: total_found counts what passes them.
							// With attribute sort keys the key does not depend on the weight: compare it whole, rank only what may enter.
							uint32_t uRows = 0;
							#pragma unroll
							for ( int k=0; k<CHUNK_K; ++k )
								if ( ( dF[k>>1]>>( ( k & 1 ) ? 24 : 8 ) ) & 0xFu )
									uRows |= 1u<<k;
							uint32_t uKeep = 0, uPass = 0;
							uMyAttrRows += __popc ( uRows );
							for ( uint32_t m=uRows; m; m&=m-1 )
							{
								const int k = __ffs ( m )-1;
								if ( !PassFilters ( tIdx, q, uRowC+k ) )
									continue;
								uKeep |= 1u<<k;
								if ( q.m_nSortKeys ? !KeyLess ( MakeKey ( tIdx, q, uRowC+k, 0x7FFFFFFF ), tThrCur ) : ( ( uCand>>k ) & 1u )!=0 )
									uPass |= 1u<<k;
							}
							uPresent = uKeep;
							uCand = uPass;
						}
						iMyTotal += __popc ( uPresent );
						if ( !__any_sync ( FULL_MASK, uCand!=0 ) )
							continue;

						// B1. candidate rows without a sparse posting join the warp's queue (row = index-local rowid); full passes run now
						{
							const uint32_t uQ = uCand & ~uSparseRows;
							int iOff = __popc ( uQ );
							#pragma unroll
							for ( int d=1; d<32; d<<=1 )
							{
								const int t = __shfl_up_sync ( FULL_MASK, iOff, d );
								if ( iLane>=d )
									iOff += t;
							}
							const int nNew = __shfl_sync ( FULL_MASK, iOff, 31 );
							if ( nNew )
							{
								iOff += nQueue - __popc ( uQ );
								for ( uint32_t m=uQ; m; m&=m-1 )
									pQueue[iOff++] = uRowC + __ffs ( m )-1;
								nQueue += nNew;
								__syncwarp();
								while ( nQueue>=32 )
								{
									nQueue -= 32;
									if constexpr ( DNF )
										fnExactDnf ( true, pQueue[nQueue+iLane], -1, nullptr, pPool, tThrCur );
									else
										fnExactHot ( true, pQueue[nQueue+iLane], pPool, tThrCur );
								}
								__syncwarp();
							}
						}

						// B2. candidate rows holding a sparse posting: compact them and evaluate now, one lane per row, all ops in order
						const uint32_t uNow = uCand & uSparseRows;
						if ( !__any_sync ( FULL_MASK, uNow!=0 ) )
							continue;
						int iOff = __popc ( uNow );
						#pragma unroll
						for ( int d=1; d<32; d<<=1 )
						{
							const int t = __shfl_up_sync ( FULL_MASK, iOff, d );
							if ( iLane>=d )
								iOff += t;
						}
						const int nCand = __shfl_sync ( FULL_MASK, iOff, 31 );
						iOff -= __popc ( uNow );
						for ( uint32_t m=uNow; m; m&=m-1 )
							pCand[iOff++] = (uint8_t)( iLane*8 + __ffs ( m )-1 );
						__syncwarp();
						for ( int iBase=0; iBase<nCand; iBase+=32 )
						{
							const bool bAct = iBase+iLane<nCand;
							const int sRow = c*256 + ( bAct ? (int)pCand[iBase+iLane] : 0 );	// slot inside the mini-tile
							if constexpr ( DNF )
								fnExactDnf ( bAct, uLo+sRow, sRow, pList, pPool, tThrCur );
							else
							{
								float fT = 0.0f;
								uint32_t uF = 0;
								bool bPres = false;
								int iSp = 0;
								for ( int iOp=0; iOp<nOps; ++iOp )
								{
									const uint16_t * pRow = S.m_dOpPtr[iOp];
									if ( pRow )
									{
										const DevLeaf_t & tLeaf = q.m_dLeaves[q.m_dOps[iOp].m_uLeaf];
										const uint32_t uRaw = bAct ? __ldg ( pRow+uLo+sRow ) : 0u;
										const uint32_t uHits = uRaw & 255u;
										const uint32_t uFields = ( uRaw>>8 ) & tLeaf.m_uQueriedFields;
										if ( !uHits || !uFields )
											continue;
										float fBase = S.m_dTf[uHits];
										if ( bAnyEscape && uHits==255 )
											fBase = HotEscapeTf ( P.m_tHot, tLeaf.m_iHot, uLo+sRow );
										const float fTf = __fmul_rn ( fBase, tLeaf.m_fIDF );
										// ExtOr_c: both sides -> sum, one side -> copy (src/searchnode.cpp:3486-3504)
										fT = bPres ? __fadd_rn ( fT, fTf ) : fTf;
										uF |= uFields;
										bPres = true;
									} else
									{
										const int iTo = S.m_dOpStart[iWarp][iSp+1];
										for ( int e=S.m_dOpStart[iWarp][iSp]; e<iTo; ++e )
										{
											const PreEntry_t tEntry = pList[e];	// same address in every lane: a broadcast
											if ( bAct && (int)tEntry.m_uRowid==sRow )
											{
												fT = bPres ? __fadd_rn ( fT, tEntry.m_fTf ) : tEntry.m_fTf;
												uF |= tEntry.m_uFields;
												bPres = true;
											}
										}
										++iSp;
									}
								}
	fnRankPush ( bAct && bPres, fT, uF, uLo+sRow, pPool, tThrCur );
							}
						}
						__syncwarp();
					}
					__syncwarp();
					++uMini;
					continue;
				}

				if constexpr ( !ORONLY )
				{
				// run the tile program on this warp's private vectors
				for ( int iOp=0; iOp<q.m_nOps; ++iOp )
				{
					const DevOp_t tOp = q.m_dOps[iOp];
					const int d = tOp.m_uDst;
					if ( tOp.m_eCode<=OP_TERM_MAYBE )
					{
						const DevLeaf_t & tLeaf = q.m_dLeaves[tOp.m_uLeaf];
						const bool bDense = tLeaf.m_iHot>=0;
						const bool bChain = ( tOp.m_uSrc!=0 );

						// AND chains: any candidate left?
						if ( tOp.m_eCode==OP_TERM_AND && bChain )
						{
							bool bMine = false;
							#pragma unroll
							for ( int k=0; k<MINI_K; ++k )
								bMine |= ( V.Cnt ( d, k*32+iLane )==tOp.m_uAliveDst );
							if ( !__any_sync ( FULL_MASK, bMine ) )
							{
								iOp = (int)tOp.m_uSrc-2;
								continue;
							}
						}

						if ( bDense )
						{
							const uint16_t * pD = P.m_tHot.m_pData + (size_t)tLeaf.m_iHot*P.m_tHot.m_iStride + uLo + iLane;
							#pragma unroll 1
							for ( int c=0; c<MINI_K/CHUNK_K; ++c )
							{
								const int iOff = c*CHUNK_K*32;
								uint32_t dRaw[CHUNK_K];
								#pragma unroll
								for ( int k=0; k<CHUNK_K; ++k )
									dRaw[k] = __ldg ( pD + iOff + k*32 );
								float * pT = V.m_pTfidf + d*MINI_W + iOff + iLane;
								uint32_t * pF = V.m_pFields + d*MINI_W + iOff + iLane;
								uint8_t * pC = V.m_pCnt + d*MINI_W + iOff + iLane;
								switch ( tOp.m_eCode )
								{
								case OP_TERM_SET:	StreamDenseOp<OP_TERM_SET> ( pT, pF, pC, dRaw, S.m_dTf, P.m_tHot, tLeaf.m_iHot, uLo+iOff+iLane, tLeaf.m_uQueriedFields, tLeaf.m_fIDF, tOp.m_uAliveDst, tOp.m_uAliveOut ); break;
								case OP_TERM_AND:	StreamDenseOp<OP_TERM_AND> ( pT, pF, pC, dRaw, S.m_dTf, P.m_tHot, tLeaf.m_iHot, uLo+iOff+iLane, tLeaf.m_uQueriedFields, tLeaf.m_fIDF, tOp.m_uAliveDst, tOp.m_uAliveOut ); break;
								case OP_TERM_OR:	StreamDenseOp<OP_TERM_OR> ( pT, pF, pC, dRaw, S.m_dTf, P.m_tHot, tLeaf.m_iHot, uLo+iOff+iLane, tLeaf.m_uQueriedFields, tLeaf.m_fIDF, tOp.m_uAliveDst, tOp.m_uAliveOut ); break;
								case OP_TERM_ANDNOT: StreamDenseOp<OP_TERM_ANDNOT> ( pT, pF, pC, dRaw, S.m_dTf, P.m_tHot, tLeaf.m_iHot, uLo+iOff+iLane, tLeaf.m_uQueriedFields, tLeaf.m_fIDF, tOp.m_uAliveDst, tOp.m_uAliveOut ); break;
								default:			StreamDenseOp<OP_TERM_MAYBE> ( pT, pF, pC, dRaw, S.m_dTf, P.m_tHot, tLeaf.m_iHot, uLo+iOff+iLane, tLeaf.m_uQueriedFields, tLeaf.m_fIDF, tOp.m_uAliveDst, tOp.m_uAliveOut ); break;
								}
							}
						} else
						{
							if ( tOp.m_eCode==OP_TERM_SET )
							{
								#pragma unroll
								for ( int k=0; k<MINI_K; ++k )
									V.Cnt ( d, k*32+iLane ) = 0;
								__syncwarp();
							}
							bool bApplied = false;
							if ( S.m_dNext[iWarp][tOp.m_uLeaf]<uHi )
							{
								const uint32_t * pBase = tIdx.m_pBlkRowid + tLeaf.m_uFirstBlk;
								uint32_t b = tLeaf.m_nBlocks ? StreamSeek ( pBase, tLeaf.m_nBlocks, S.m_dCur[iWarp][tOp.m_uLeaf], uLo, iLane ) : 0;
								uint32_t uNextRow = 0xFFFFFFFFu;	// bound of the keyword's next rowid after this mini-tile
								while ( b<tLeaf.m_nBlocks )
								{
									const uint32_t uBase = __ldg ( pBase+b );
									if ( uBase>=uHi )
									{
										uNextRow = uBase;
										break;
									}
									const uint32_t uNextBase = b+1<tLeaf.m_nBlocks ? __ldg ( pBase+b+1 ) : 0xFFFFFFFFu;
									bool bNeed = true;
									if ( tOp.m_eCode==OP_TERM_AND && bChain )
									{
										// any candidate inside this block's rowid range [base_b, base_b+1) ?
										const uint32_t uFrom = max ( uBase, uLo )-uLo, uTo = min ( uNextBase, uHi )-uLo;
										bool bAny = false;
										for ( uint32_t uS=( uFrom & ~31u )+iLane; uS<uTo; uS+=32 )
											bAny |= ( uS>=uFrom && V.Cnt ( d, (int)uS )==tOp.m_uAliveDst );
										bNeed = __any_sync ( FULL_MASK, bAny );
									}
									if ( bNeed )
									{
										StreamCacheBlock ( tIdx, tLeaf, b, &S.m_dCached[iWarp][tOp.m_uLeaf], pCache0+tOp.m_uLeaf*32, pStage, pRecStart, S.m_dTf, iLane );
										const PreEntry_t tEntry = pCache0[tOp.m_uLeaf*32+iLane];
										if ( tEntry.m_uRowid>=uLo && tEntry.m_uRowid<uHi )
										{
											ApplyTermOp<false> ( V, tOp, d, (int)( tEntry.m_uRowid-uLo ), tEntry.m_fTf, tEntry.m_uFields, 0u );
											bApplied = true;
										}
									}
									if ( uNextBase>uHi )
									{
										// the block's last rowid (next base - 1) lies beyond this mini-tile: keep it, and learn its next rowid
										uNextRow = uHi;
										if ( S.m_dCached[iWarp][tOp.m_uLeaf]==b )
										{
											const uint32_t r = pCache0[tOp.m_uLeaf*32+iLane].m_uRowid;
											uint32_t uMin = ( r!=0xFFFFFFFFu && r>=uHi ) ? r : 0xFFFFFFFFu;
											#pragma unroll
											for ( int iStep=16; iStep; iStep>>=1 )
												uMin = min ( uMin, __shfl_xor_sync ( FULL_MASK, uMin, iStep ) );
											uNextRow = min ( uMin, uNextBase );
										}
										break;
									}
									++b;
								}
								__syncwarp();
								if ( iLane==0 )
								{
									S.m_dCur[iWarp][tOp.m_uLeaf] = b;
									S.m_dNext[iWarp][tOp.m_uLeaf] = uNextRow;
								}
							}
							__syncwarp();
							if ( tOp.m_eCode==OP_TERM_SET && bChain && !__any_sync ( FULL_MASK, bApplied ) )
							{
								iOp = (int)tOp.m_uSrc-2;	// an opening keyword without postings here: the whole chain is empty
								continue;
							}
						}
						__syncwarp();
					} else if ( tOp.m_eCode!=OP_NWAY )
					{
						const int r = tOp.m_uSrc;
						const uint8_t uAd = tOp.m_uAliveDst, uAs = tOp.m_uAliveSrc;
						#pragma unroll 4
						for ( int k=0; k<MINI_K; ++k )
						{
							const int s = k*32+iLane;
							const bool bD = V.Cnt ( d, s )==uAd, bS = V.Cnt ( r, s )==uAs;
							switch ( tOp.m_eCode )
							{
							case OP_VEC_AND:
								if ( bD && bS )
								{
									V.Tfidf ( d, s ) = __fadd_rn ( V.Tfidf ( d, s ), V.Tfidf ( r, s ) );
									V.Fields ( d, s ) |= V.Fields ( r, s );
								} else if ( bD )
									V.Cnt ( d, s ) = 0;
								break;
							case OP_VEC_OR:
								if ( bD && bS )
								{
									V.Tfidf ( d, s ) = __fadd_rn ( V.Tfidf ( d, s ), V.Tfidf ( r, s ) );
									V.Fields ( d, s ) |= V.Fields ( r, s );
								} else if ( bS )
								{
									V.Tfidf ( d, s ) = V.Tfidf ( r, s ); V.Fields ( d, s ) = V.Fields ( r, s ); V.Cnt ( d, s ) = uAd;
								}
								break;
							case OP_VEC_ANDNOT:
								if ( bD && bS )
									V.Cnt ( d, s ) = 0;
								break;
							case OP_VEC_MAYBE:
								if ( bD && bS )
								{
									V.Tfidf ( d, s ) = __fadd_rn ( V.Tfidf ( d, s ), V.Tfidf ( r, s ) );
									V.Fields ( d, s ) |= V.Fields ( r, s );
								}
								break;
							default:
								break;
							}
						}
						__syncwarp();
					}
				}

				// rank + filter + push survivors of this mini-tile: 8 rows per lane
				const int nValid = (int)( uHi-uLo );
				if ( bFastRank )
				{
					// the common shape (BM25 weights, relevance order, no filters): a 32-bit weight compare rejects most rows
					const uint32_t uThrWx = (uint32_t)( tThr.m_uHi>>32 ), uThrRow = ~(uint32_t)( tThr.m_uLo>>32 );
					const float * pT = V.m_pTfidf + iLane;
					const uint32_t * pF = V.m_pFields + iLane;
					const uint8_t * pC = V.m_pCnt + iLane;
					#pragma unroll 8
					for ( int k=0; k<MINI_K; ++k )
					{
						bool bPush = false;
						uint32_t uW = 0;
						if ( k*32+iLane<nValid && pC[k*32]==uAliveRoot )
						{
							// seed weight src/sphinxsearch.cpp:1070, ExtRanker_WeightSum_c :1112-1129
							const int iSeed = __float2int_rz ( __fmul_rn ( __fadd_rn ( pT[k*32], 0.5f ), 1000.0f ) );
							const uint32_t uMask = pF[k*32];
							const uint32_t uRank = uMask ? S.m_dRankTab[uMask & 15u] : 1u;
							uW = ( (uint32_t)iSeed + uRank*1000u )*(uint32_t)q.m_iIndexWeight;
							++iMyTotal;
							const uint32_t uWx = uW ^ 0x80000000u;
							bPush = uWx>uThrWx || ( uWx==uThrWx && uLo+k*32+iLane+tIdx.m_uRowidBase<=uThrRow );
						}
						const unsigned m = __ballot_sync ( FULL_MASK, bPush );
						if ( m )
						{
							int iBase = 0;
							if ( iLane==0 )
								iBase = atomicAdd ( &S.m_iPoolCnt, __popc ( m ) );
							iBase = __shfl_sync ( FULL_MASK, iBase, 0 );
							if ( bPush )
							{
								Key128_t tKey;
								tKey.m_uHi = (uint64_t)( uW ^ 0x80000000u )<<32;
								tKey.m_uLo = ( (uint64_t)( ~( uLo+k*32+iLane+tIdx.m_uRowidBase ) )<<32 ) | uW;
								pPool[iBase + __popc ( m & ( ( 1u<<iLane )-1u ) )] = tKey;
							}
						}
					}
				} else
				#pragma unroll 2
				for ( int k=0; k<MINI_K; ++k )
				{
					const int s = k*32+iLane;
					bool bPush = false;
					Key128_t tKey;
					if ( s<nValid && V.Cnt ( 0, s )==uAliveRoot )
					{
						const uint32_t uRowid = uLo+s;
						bool bOk = PassFilters ( tIdx, q, uRowid );
						int iWeight = 1;	// ExtRanker_None_c, src/sphinxsearch.cpp:1160
						if ( bOk && q.m_eRanker!=2 )
						{
							// seed weight src/sphinxsearch.cpp:1070, ExtRanker_WeightSum_c :1112-1129
							const int iSeed = __float2int_rz ( __fmul_rn ( __fadd_rn ( V.Tfidf ( 0, s ), 0.5f ), 1000.0f ) );
							const uint32_t uMask = V.Fields ( 0, s );
							uint32_t uRank = 0;
							if ( !uMask )
								uRank = 1;
							else if ( q.m_nWeights<=4 )
								uRank = S.m_dRankTab[uMask & 15u];
							else
								for ( int i=0; i<q.m_nWeights; ++i )
									if ( uMask & ( 1u<<i ) )
										uRank += (uint32_t)q.m_dWeights[i];
							iWeight = q.m_eRanker==4 ? (int)uRank : (int)( (uint32_t)iSeed + uRank*1000u );	// ExtRanker_WeightSum_c<false> under SPH_RANK_PROXIMITY
						}
						if ( bOk && tIdx.m_pDead )
							bOk = !( ( __ldg ( tIdx.m_pDead+( uRowid>>5 ) )>>( uRowid & 31 ) ) & 1u );
						if ( bOk )
						{
							iWeight = (int)( (uint32_t)iWeight*(uint32_t)q.m_iIndexWeight );
							++iMyTotal;
							tKey = MakeKey ( tIdx, q, uRowid, iWeight );
							bPush = !KeyLess ( tKey, tThr );
						}
					}
					const unsigned m = __ballot_sync ( FULL_MASK, bPush );
					if ( m )
					{
						int iBase = 0;
						if ( iLane==0 )
							iBase = atomicAdd ( &S.m_iPoolCnt, __popc ( m ) );
						iBase = __shfl_sync ( FULL_MASK, iBase, 0 );
						if ( bPush )
							pPool[iBase + __popc ( m & ( ( 1u<<iLane )-1u ) )] = tKey;
					}
				}
				__syncwarp();
				++uMini;
							}
			}
			// register-OR path: the queued candidate rows are evaluated against this round's threshold and pool buffer
			if ( nQueue )
			{
				if constexpr ( DNF )
					fnExactDnf ( iLane<nQueue, pQueue[iLane<nQueue ? iLane : 0], -1, nullptr, pPool, tThrCur );
				else
					fnExactHot ( iLane<nQueue, pQueue[iLane<nQueue ? iLane : 0], pPool, tThrCur );
				nQueue = 0;
				__syncwarp();
			}
		}
		__syncthreads();

		// item epilogue: final selection, publish keys + counters
		if ( S.m_iPoolCnt>iK )
		{
			Key128_t * pIn = pPool0 + (size_t)S.m_iPoolBuf*P.m_iPoolCap;
			Key128_t * pOut = pPool0 + (size_t)( S.m_iPoolBuf^1 )*P.m_iPoolCap;
			CtaSelectTopK ( pIn, S.m_iPoolCnt, iK, pOut, S.m_tSel );
			if ( tid==0 )
			{
				S.m_iPoolCnt = iK;
				S.m_iPoolBuf ^= 1;
			}
			__syncthreads();
		}
		{
			#pragma unroll
			for ( int d=16; d; d>>=1 )
				iMyTotal += __shfl_xor_sync ( FULL_MASK, iMyTotal, d );
			if ( iLane==0 && iMyTotal )
				atomicAdd ( &S.m_uTotal, (unsigned long long)iMyTotal );
			if ( DNF && P.m_pWork )
			{
				#pragma unroll
				for ( int d=16; d; d>>=1 )
					uMyAttrRows += __shfl_xor_sync ( FULL_MASK, uMyAttrRows, d );
				if ( iLane==0 && uMyAttrRows )
					atomicAdd ( P.m_pWork+1, (unsigned long long)uMyAttrRows );
			}
			__syncthreads();
			const Key128_t * pPool = pPool0 + (size_t)S.m_iPoolBuf*P.m_iPoolCap;
			Key128_t * pDst = P.m_pItemKeys + (size_t)iItem*P.m_iKMax;
			const int n = S.m_iPoolCnt;
			for ( int i=tid; i<n; i+=EVAL_THREADS )
				pDst[i] = pPool[i];
			if ( tid==0 )
			{
				P.m_pItemOut[iItem].m_iTotalFound = (int64_t)S.m_uTotal;
				P.m_pItemOut[iItem].m_nKeys = n;
			}
		}
	}
}

