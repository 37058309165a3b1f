// K2/K3/K6/K7/K8 for doc-only queries: the warp-autonomous streaming evaluator.
//
// A work item = (query, rowid range). The CTA cuts the range into 8 contiguous sub-ranges, one per warp; every warp walks
// its sub-range in 256-row mini-tiles with private dense vectors in shared memory and NO CTA-wide barrier inside the
// tile program, so that an SM has 30-60 independent latency chains in flight instead of one per CTA:
//   - hot keywords: 8 rows per lane straight from the batch's dense store (u16 per row);
//   - sparse keywords: a per-(warp, keyword) cursor over the resident block table; the current 32-doc block is decoded once
//     per warp into an L1/L2-resident cache and re-used by all the mini-tiles it spans; AND chains test the candidates'
//     bitmap before decoding a block (DiskIndexQword_c::HintRowID skip, src/sphinx.cpp:407-451);
//   - mini-tiles that no document-originating keyword touches are jumped over using the exact next rowid;
//   - survivors are ranked per lane and pushed into the CTA's candidate pool; warps meet at one barrier every 8 mini-tiles,
//     where the pool is compacted to K by the CTA radix select if it could overflow.
// Semantics per op are those of eval_kernel (ApplyTermOp): ExtTerm_T / ExtAnd_c / ExtMultiAnd_T / ExtOr_c / ExtAndNot_c /
// ExtMaybe_c of src/searchnode.cpp, TF*IDF in the reference's association order.
#pragma once
// (included from kernels.cu inside namespace mgpu: uses its DecodeBlock / ApplyTermOp / CtaSelectTopK / MakeKey helpers)

static const int OR_LIST_CAP = 512*MAX_LEAVES;	///< register-OR path: a mini-tile holds at most 512 postings of each of <=16 sparse keywords
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

template<int MINI_W>
__global__ void __launch_bounds__ ( EVAL_THREADS, 4 ) stream_kernel ( EvalParams_t P, int nStack )
{
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
			S.m_iItem = atomicAdd ( P.m_pCounter, 1 );
		__syncthreads();
		const int iItem = S.m_iItem;
		if ( iItem>=P.m_nItems )
			break;
		const DevWorkItem_t tItem = P.m_pItems[iItem];
		{
			const uint32_t * pSrc = reinterpret_cast<const uint32_t *>( P.m_pQueries+tItem.m_uQuery );
			uint32_t * pDst = reinterpret_cast<uint32_t *>( &S.m_tQ );
			for ( int i=tid; i<(int)( sizeof(DevQuery_t)/4 ); i+=EVAL_THREADS )
				pDst[i] = pSrc[i];
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
		}
		int iMyTotal = 0;

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
		const bool bRegOr = bFastRank && q.m_bPureOr && P.m_pOrList;
		const bool bAnyEscape = P.m_tHot.m_pEscapeCount && __ldg ( P.m_tHot.m_pEscapeCount )!=0;	// documents with >= 255 hits of a hot keyword exist

		for ( uint32_t uRound=0; uRound<nRounds; ++uRound )
		{
			// all warps meet here; compact the candidate pool if this round could overflow it
			__syncthreads();
			const int iPoolNow = S.m_iPoolCnt;
			__syncthreads();	// nobody pushes before everybody has read the level
			if ( iPoolNow+STREAM_POOL_SLACK>P.m_iPoolCap || ( uRound<=2 && iPoolNow>iK ) )
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
			const uint32_t uRoundEnd = min ( uMini1, uMini0 + ( uRound<2 ? uRound+1 : 2+( uRound-1 )*SYNC_MINIS ) );

			while ( uMini<uRoundEnd )
			{
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
				if ( MINI_W==512 && bRegOr )
				{
					PreEntry_t * pList = P.m_pOrList + ( (size_t)blockIdx.x*EVAL_WARPS+iWarp )*OR_LIST_CAP;
					int nList = 0;
					for ( int iOp=0; iOp<q.m_nOps; ++iOp )
					{
						const int l = q.m_dOps[iOp].m_uLeaf;
						const DevLeaf_t & tLeaf = q.m_dLeaves[l];
						if ( iLane==0 )
							S.m_dOpStart[iWarp][iOp] = (uint16_t)nList;
						if ( tLeaf.m_iHot>=0 || S.m_dNext[iWarp][l]>=uHi )
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
							const bool bIn = tEntry.m_uRowid>=uLo && tEntry.m_uRowid<uHi;
							const unsigned m = __ballot_sync ( FULL_MASK, bIn );
							if ( bIn )
							{
								tEntry.m_uRowid -= uLo;	// slot inside the mini-tile
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
						S.m_dOpStart[iWarp][q.m_nOps] = (uint16_t)nList;
					__syncwarp();

					const uint32_t uThrWx = (uint32_t)( tThr.m_uHi>>32 ), uThrRow = ~(uint32_t)( tThr.m_uLo>>32 );
					const int nValid = (int)( uHi-uLo );
					#pragma unroll 1
					for ( int c=0; c<MINI_W/32/CHUNK_K; ++c )
					{
						float dT[CHUNK_K];
						uint32_t dF[CHUNK_K];
						uint32_t uPres = 0;
						#pragma unroll
						for ( int k=0; k<CHUNK_K; ++k )
						{
							dT[k] = 0.0f; dF[k] = 0;
						}
						const int iRow0 = c*CHUNK_K*32 + iLane;	// this lane's rows: iRow0 + 32k
						for ( int iOp=0; iOp<q.m_nOps; ++iOp )
						{
							const DevLeaf_t & tLeaf = q.m_dLeaves[q.m_dOps[iOp].m_uLeaf];
							if ( tLeaf.m_iHot>=0 )
							{
								const uint16_t * pD = P.m_tHot.m_pData + (size_t)tLeaf.m_iHot*P.m_tHot.m_iStride + uLo + iRow0;
								uint32_t dRaw[CHUNK_K];
								#pragma unroll
								for ( int k=0; k<CHUNK_K; ++k )
									dRaw[k] = __ldg ( pD + k*32 );
								const uint32_t uQueried = tLeaf.m_uQueriedFields;
								const float fIDF = tLeaf.m_fIDF;
								#pragma unroll
								for ( int k=0; k<CHUNK_K; ++k )
								{
									if ( !dRaw[k] )
										continue;	// keyword absent from this row (a present row has hits >= 1)
									const uint32_t uHits = dRaw[k] & 255u;
									const uint32_t uFields = ( dRaw[k]>>8 ) & uQueried;
									if ( !uFields )
										continue;
									float fBase = S.m_dTf[uHits];
									if ( bAnyEscape && uHits==255 )
										fBase = HotEscapeTf ( P.m_tHot, tLeaf.m_iHot, uLo+iRow0+k*32 );
									const float fTf = __fmul_rn ( fBase, fIDF );
									// ExtOr_c: both sides -> sum, one side -> copy (src/searchnode.cpp:3486-3504)
									dT[k] = ( uPres>>k ) & 1u ? __fadd_rn ( dT[k], fTf ) : fTf;
									dF[k] |= uFields;
									uPres |= 1u<<k;
								}
							} else
							{
								const int iTo = S.m_dOpStart[iWarp][iOp+1];
								for ( int e=S.m_dOpStart[iWarp][iOp]; e<iTo; ++e )
								{
									const PreEntry_t tEntry = pList[e];	// same address in every lane: a broadcast
									const int sRow = (int)tEntry.m_uRowid;
									if ( ( sRow>>8 )!=c || ( sRow & 31 )!=iLane )
										continue;
									const int kk = ( sRow>>5 ) & ( CHUNK_K-1 );
									#pragma unroll
									for ( int k=0; k<CHUNK_K; ++k )
										if ( k==kk )
										{
											dT[k] = ( uPres>>k ) & 1u ? __fadd_rn ( dT[k], tEntry.m_fTf ) : tEntry.m_fTf;
											dF[k] |= tEntry.m_uFields;
											uPres |= 1u<<k;
										}
								}
							}
						}

						// rank + threshold, straight from the registers; rows that beat the K-th best so far are rare after warm-up, so the
						// warp votes once per chunk and only then walks the rows to push
						uint32_t uValid = uPres;
						#pragma unroll
						for ( int k=0; k<CHUNK_K; ++k )
							if ( iRow0+k*32>=nValid )
								uValid &= ~( 1u<<k );
						iMyTotal += __popc ( uValid );
						uint32_t dW[CHUNK_K];
						uint32_t uPush = 0;
						#pragma unroll
						for ( int k=0; k<CHUNK_K; ++k )
						{
							// seed weight src/sphinxsearch.cpp:1070, ExtRanker_WeightSum_c :1112-1129
							const int iSeed = __float2int_rz ( __fmul_rn ( __fadd_rn ( dT[k], 0.5f ), 1000.0f ) );
							const uint32_t uRank = dF[k] ? S.m_dRankTab[dF[k] & 15u] : 1u;
							dW[k] = ( (uint32_t)iSeed + uRank*1000u )*(uint32_t)q.m_iIndexWeight;
							const uint32_t uWx = dW[k] ^ 0x80000000u;
							if ( ( ( uValid>>k ) & 1u ) && ( uWx>uThrWx || ( uWx==uThrWx && uLo+iRow0+k*32+tIdx.m_uRowidBase<=uThrRow ) ) )
								uPush |= 1u<<k;
						}
						if ( __any_sync ( FULL_MASK, uPush!=0 ) )
						{
							#pragma unroll
							for ( int k=0; k<CHUNK_K; ++k )
							{
								const bool bPush = ( uPush>>k ) & 1u;
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
										tKey.m_uHi = (uint64_t)( dW[k] ^ 0x80000000u )<<32;
										tKey.m_uLo = ( (uint64_t)( ~( uLo+iRow0+k*32+tIdx.m_uRowidBase ) )<<32 ) | dW[k];
										pPool[iBase + __popc ( m & ( ( 1u<<iLane )-1u ) )] = tKey;
									}
								}
							}
						}
					}
					__syncwarp();
					++uMini;
					continue;
				}

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

