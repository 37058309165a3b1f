// K3c orbits_kernel: launch class 5, pure OR programs under BM25 in relevance order ("OR over bit planes").
//
// A top-K OR query has to COUNT every matching row (total_found = every push, src/sphinxsort.cpp:724) but only has to RANK the
// rows that can still enter the top K. The batch's hot-term store carries, next to the u16 {hits, fields, tf class} per row, one
// presence BITMAP per (hot keyword, field): 1 bit per row instead of 16. A warp walks its share of the work item in 1024-row
// mini-tiles, one 32-bit word (32 rows) per lane and bitmap:
//
//  1. bitmap pass, ~1 instruction per 32 rows and keyword: per-field ORs give the matched-field mask F(row) bit-parallel, the
//     per-keyword presence words P_i go to shared memory; popc ( present ) is the exact total_found.
//  2. candidate selection, MaxScore style (Turtle & Flood), bit-parallel: a row's weight is at most
//     rank ( F )*1000 + 500 + 1000*sum over the PRESENT keywords of max ( idf_i, 0 ) (tf < 1; ExtRanker_WeightSum_c,
//     src/sphinxsearch.cpp:1096-1141). For every rank class fv (value of F) the warp knows how much TF*IDF a row of that class needs
//     to reach the K-th best weight so far (T); with the hot keywords sorted by their bound, the rows that can still make it are
//     those holding ALL keywords of a prefix ("required": without one of them the rest cannot reach T) or, if none is required,
//     ANY keyword of the "essential" prefix (the rest together stays below T). Rows holding a sparse keyword's posting always
//     qualify while their class can reach the threshold at all.
//     Keywords with a NEGATIVE idf (df > N/2: the stop words) lower the weight of nearly every row by about as much as the
//     rare keywords raise it, so ignoring them leaves ~20 % of the rows as candidates (scripts/or_bound_study.py). Up to two of
//     them refine the classes by their tf level on the row (absent / 1 hit / 2-3 hits / >= 4 hits, from two more bitmaps the
//     store keeps for keywords in >= 1/3 of the rows): a row of level l owes at least tf_min ( l )*|idf|, which raises the
//     TF*IDF its positive keywords must bring. The reachable (rank class, level, level) combinations are listed whenever the
//     warp's threshold moves; ~1 % of the rows stay candidates.
//  3. exact pass for the candidate rows only, one lane per row, TF*IDF in op order (ExtOr_c left fold,
//     src/searchnode.cpp:3486-3504) from the u16 store / the mini-tile's list of sparse postings, ranked and pushed as in K3.
//
// Everything around it (work items, sparse keyword cursors, candidate pool, on-demand CTA radix select, the query's shared
// K-th-best bound) is that of stream_kernel<512,1>, which stays available for A/B runs (mgpu_index_set_option "or_bits" = 0).
#pragma once
// (included from kernels.cu inside namespace mgpu, after stream_kernel.cuh)

static const int OB_MINI = 1024;					///< rows per warp step: one bitmap word per lane
static const int OB_LIST_CAP = OB_MINI*MAX_LEAVES;	///< sparse postings of one mini-tile
static const int OB_QUEUE = 32 + OB_MINI;			///< per-warp queue of candidate rows (its free tail doubles as the compaction scratch)
static const size_t OB_WARP_SMEM = ( MAX_LEAVES*32 + OB_QUEUE )*4;
static const int OB_MAX_CLASSES = 96;				///< reachable (rank class, level, level) combinations listed per warp; more -> levels are ignored

struct OrBitsShared_t
{
	DevQuery_t		m_tQ;
	SelectSmem_t	m_tSel;
	Key128_t		m_tThr;
	unsigned long long m_uTotal;
	int				m_iItem;
	int				m_iPoolCnt;
	int				m_iPoolBuf;
	uint32_t		m_dRankTab[16];
	int32_t			m_dRankUb[16];						///< ( field-weight sum*1000 + 500 )*64 + rounding margin, per matched-field mask
	int32_t			m_bBound;
	// bitmap pass: hot keywords sorted by their weight bound, biggest first
	const uint32_t * m_dBitPtr[MAX_LEAVES];				///< the keyword's field-0 bitmap (field f: + f*bit stride)
	uint32_t		m_dBitFields[MAX_LEAVES];			///< queried fields the index has
	int32_t			m_dUb[MAX_LEAVES];					///< ceil ( max ( idf, 0 )*64000 ) + 1
	int32_t			m_dSuffix[MAX_LEAVES+1];			///< sum of m_dUb[i..]
	uint8_t			m_dSortLeaf[MAX_LEAVES];			///< the sorted entry's leaf
	int32_t			m_nHot;
	int32_t			m_iUbSparse;						///< the same bound summed over the sparse keywords
	// penalty classes: up to two hot keywords with idf < 0 that own tf-level bitmaps
	int32_t			m_nNeg;
	int32_t			m_dNegPsm[2];						///< the keyword's entry in the sorted list (its presence word)
	const uint32_t * m_dNegLvl[2];						///< its ">= 2 hits" bitmap (">= 4 hits": + bit stride)
	int32_t			m_dNegPen[2][4];					///< what a row of tf level l owes at least, in the bound's fixed point
	uint32_t		m_dClass[EVAL_WARPS][OB_MAX_CLASSES];	///< fv | l1<<4 | l2<<7 | mode<<10 | prefix length<<16 (level 4 = any)
	// exact pass: op order
	const uint16_t * m_dOpPtr[MAX_LEAVES];				///< the op's row of the u16 store (null = sparse keyword)
	uint8_t			m_dHotLeaf[MAX_LEAVES];
	const uint16_t * m_dHotPtr[MAX_LEAVES+4];
	int32_t			m_nHotOps;
	uint8_t			m_dSparseOp[MAX_LEAVES];
	int32_t			m_nSparseOps;
	float			m_dTf[256];
	uint32_t		m_dCur[EVAL_WARPS][MAX_LEAVES];
	uint32_t		m_dCached[EVAL_WARPS][MAX_LEAVES];
	uint16_t		m_dOpStart[EVAL_WARPS][MAX_LEAVES+2];
	uint32_t		m_dNext[EVAL_WARPS][MAX_LEAVES];
	uint32_t		m_dOv[EVAL_WARPS][5][32];			///< sparse postings of the mini-tile: per-field bitmaps [0..3], rows holding any [4]
	uint16_t		m_dRecStart[EVAL_WARPS][34];
	__align__(16) uint8_t m_dStage[EVAL_WARPS][STAGE_BYTES];
};

__global__ void __launch_bounds__ ( EVAL_THREADS, 3 ) orbits_kernel ( EvalParams_t P )
{
	extern __shared__ __align__(16) uint8_t dDyn[];
	__shared__ OrBitsShared_t S;
	const int tid = threadIdx.x, iWarp = tid>>5, iLane = tid & 31;
	const DevIndex_t & tIdx = P.m_tIndex;

	uint32_t * pPsm = reinterpret_cast<uint32_t *>( dDyn + (size_t)iWarp*OB_WARP_SMEM );	// [MAX_LEAVES][32] presence words of the hot keywords
	uint32_t * pQueue = pPsm + MAX_LEAVES*32;												// [OB_QUEUE]
	Key128_t * pPool0 = P.m_pPool + (size_t)blockIdx.x*2*P.m_iPoolCap;
	PreEntry_t * pCache0 = P.m_pPre + ( (size_t)blockIdx.x*EVAL_WARPS+iWarp )*MAX_LEAVES*32;
	PreEntry_t * pList = P.m_pOrList + ( (size_t)blockIdx.x*EVAL_WARPS+iWarp )*OB_LIST_CAP;
	uint8_t * pStage = S.m_dStage[iWarp];
	uint16_t * pRecStart = S.m_dRecStart[iWarp];
	uint32_t ( &dOv )[5][32] = S.m_dOv[iWarp];
	const size_t iBitStride = (size_t)P.m_tHot.m_iBitStride;
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
			// Is the integer weight bound usable? ( seed + field-weight sum*1000 )*index weight must stay far from 2^31 and every idf
			// must be a sane number (|idf| <= 0.5/keywords for any sane collection statistics). If not, every present row is ranked.
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
			bOk = bOk && fSum<1.0f;
			S.m_dRankUb[tid] = bOk ? (int32_t)( ( uSum*1000u + 500u )*64u + 2u ) : 0;
			if ( !tid )
				S.m_bBound = bOk ? 1 : 0;
		}
		if ( tid==32 )
		{
			int nHot = 0, nHotOps = 0, nSparse = 0, iUbSparse = 0;
			const uint32_t uIndexFields = ( 1u<<P.m_tHot.m_nBitFields )-1u;
			for ( int iOp=0; iOp<q.m_nOps && iOp<MAX_LEAVES; ++iOp )
			{
				const int l = q.m_dOps[iOp].m_uLeaf;
				const DevLeaf_t & tLeaf = q.m_dLeaves[l];
				const float fIDF = tLeaf.m_fIDF;
				// tf < 1: the keyword adds less than idf when it sits on the row, and nothing above 0 when idf <= 0
				const int iUb = fIDF>0.0f ? (int)ceilf ( __fmul_rn ( fminf ( fIDF, 1.0f ), 64000.0f ) )+1 : 0;
				S.m_dOpPtr[iOp] = nullptr;
				if ( tLeaf.m_iHot<0 )
				{
					S.m_dSparseOp[nSparse++] = (uint8_t)iOp;
					iUbSparse += iUb;
					continue;
				}
				const uint16_t * pRow = P.m_tHot.m_pData + (size_t)tLeaf.m_iHot*P.m_tHot.m_iStride;
				S.m_dOpPtr[iOp] = pRow;
				S.m_dHotPtr[nHotOps] = pRow;
				S.m_dHotLeaf[nHotOps++] = (uint8_t)l;
				// insertion into the list sorted by bound, biggest first
				int j = nHot++;
				for ( ; j>0 && S.m_dUb[j-1]<iUb; --j )
				{
					S.m_dUb[j] = S.m_dUb[j-1];
					S.m_dBitPtr[j] = S.m_dBitPtr[j-1];
					S.m_dBitFields[j] = S.m_dBitFields[j-1];
					S.m_dSortLeaf[j] = S.m_dSortLeaf[j-1];
				}
				S.m_dSortLeaf[j] = (uint8_t)l;
				S.m_dUb[j] = iUb;
				S.m_dBitPtr[j] = P.m_tHot.m_pBits + (size_t)tLeaf.m_iHot*P.m_tHot.m_nBitFields*iBitStride;
				S.m_dBitFields[j] = tLeaf.m_uQueriedFields & uIndexFields;
			}
			int iSum = 0;
			S.m_dSuffix[nHot] = 0;
			for ( int i=nHot-1; i>=0; --i )
			{
				iSum += S.m_dUb[i];
				S.m_dSuffix[i] = iSum;
			}
			// the two most negative keywords that own tf-level bitmaps
			int nNeg = 0;
			float dNegIdf[2] = { 0.0f, 0.0f };
			for ( int i=0; i<nHot; ++i )
			{
				const DevLeaf_t & tLeaf = q.m_dLeaves[S.m_dSortLeaf[i]];
				const int iLvl = ( P.m_tHot.m_pLvlSlot && tLeaf.m_fIDF<0.0f ) ? __ldg ( P.m_tHot.m_pLvlSlot+tLeaf.m_iHot ) : -1;
				if ( iLvl<0 )
					continue;
				int k = nNeg<2 ? nNeg : ( tLeaf.m_fIDF<dNegIdf[0] || tLeaf.m_fIDF<dNegIdf[1] ) ? ( dNegIdf[0]>dNegIdf[1] ? 0 : 1 ) : -1;
				if ( k<0 )
					continue;
				dNegIdf[k] = tLeaf.m_fIDF;
				S.m_dNegPsm[k] = i;
				S.m_dNegLvl[k] = P.m_tHot.m_pLvlBits + (size_t)iLvl*2*iBitStride;
				// tf >= 1/2.2, 2/3.2, 4/5.2 for >= 1, 2, 4 hits (constants rounded down, the product too)
				const float fAbs = __fmul_rn ( fminf ( -tLeaf.m_fIDF, 1.0f ), 64000.0f );
				S.m_dNegPen[k][0] = 0;
				S.m_dNegPen[k][1] = max ( 0, (int)floorf ( __fmul_rn ( fAbs, 0.4545f ) )-1 );
				S.m_dNegPen[k][2] = max ( 0, (int)floorf ( __fmul_rn ( fAbs, 0.6249f ) )-1 );
				S.m_dNegPen[k][3] = max ( 0, (int)floorf ( __fmul_rn ( fAbs, 0.7692f ) )-1 );
				if ( nNeg<2 )
					++nNeg;
			}
			S.m_nNeg = nNeg;
			S.m_nHot = nHot;
			S.m_nHotOps = nHotOps;
			S.m_nSparseOps = nSparse;
			S.m_iUbSparse = iUbSparse;
		}
		int iMyTotal = 0;
		uint32_t uDbgMinis = 0, uDbgHot = 0, uDbgSparse = 0;

		// this warp's contiguous share of the item
		const uint32_t nMinis = ( tItem.m_uRowHi-tItem.m_uRowLo+OB_MINI-1 )/OB_MINI;
		const uint32_t uMini0 = (uint32_t)( (uint64_t)nMinis*iWarp/EVAL_WARPS ), uMini1 = (uint32_t)( (uint64_t)nMinis*( iWarp+1 )/EVAL_WARPS );
		uint32_t uMini = uMini0;
		const bool bAnyEscape = P.m_tHot.m_pEscapeCount && __ldg ( P.m_tHot.m_pEscapeCount )!=0;

		__syncthreads();	// S.m_nHot and friends
		int nQueue = 0;
		uint32_t uNextSparse = 0;
		const int nOps = q.m_nOps, nHot = S.m_nHot, nHotOps = S.m_nHotOps, nSparseOps = S.m_nSparseOps;
		const int iUbHot = S.m_dSuffix[0], iUbSparse = S.m_iUbSparse;
		const int iIndexWeight = q.m_iIndexWeight;
		const int nFv = 1<<P.m_tHot.m_nBitFields;
		const int nNeg = S.m_nNeg;
		uint32_t * pClass = S.m_dClass[iWarp];

		// ranks one evaluated row and pushes it if it beats the K-th best key so far (one row per lane)
		auto fnRankPush = [&] ( bool bRow, float fT, uint32_t uF, uint32_t uRow, Key128_t * pPool, const Key128_t & tThr )
		{
			// seed weight src/sphinxsearch.cpp:1070, ExtRanker_WeightSum_c :1112-1129
			const int iSeed = __float2int_rz ( __fmul_rn ( __fadd_rn ( fT, 0.5f ), 1000.0f ) );
			const uint32_t uRank = uF ? S.m_dRankTab[uF & 15u] : 1u;
			const uint32_t uW = ( (uint32_t)iSeed + uRank*1000u )*(uint32_t)iIndexWeight;
			Key128_t tKey;
			tKey.m_uHi = (uint64_t)( uW ^ 0x80000000u )<<32;
			tKey.m_uLo = ( (uint64_t)( ~( uRow+tIdx.m_uRowidBase ) )<<32 ) | uW;
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

		// The pool is compacted whenever it holds iTrigger keys: every warp looks at the level before each mini-tile and comes to the
		// barrier when it is reached (a warp adds at most 1024+31 keys in between, far below the pool's slack).
		const int iTrigger = min ( 2*iK+1024, iK+16384 );
		while ( true )
		{
			// all warps meet here; compact the candidate pool if it reached the trigger
			__syncthreads();
			const int iPoolNow = S.m_iPoolCnt;
			if ( !__syncthreads_or ( uMini<uMini1 ) )
				break;	// every warp has finished its share
			if ( iPoolNow>=iTrigger )
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
			Key128_t tThrCur;	// the warp's current bound of the K-th best key
			int nClasses = 0;	// reachable (rank class, tf level of negative keyword 1, of negative keyword 2) combinations in pClass
			// how the rows of a class are selected given the TF*IDF its positive keywords must bring: mode 0 = no row can reach the
			// threshold, 1 = only rows holding a sparse posting, 2 = every row, 3 = rows holding all of the first n hot keywords,
			// 4 = rows holding any of the first n hot keywords (3, 4: or a sparse posting)
			auto fnMode = [&] ( int iNeed ) -> uint32_t
			{
				if ( iNeed<=0 )
					return 2u<<10;
				if ( iNeed>iUbHot+iUbSparse )
					return 0u;
				if ( iNeed>iUbHot )
					return 1u<<10;
				int r = 0;
				while ( r<nHot && S.m_dUb[r]>iUbHot-iNeed )
					++r;
				if ( r>0 )
					return ( 3u<<10 ) | ( (uint32_t)r<<16 );
				int e = 1;
				while ( e<nHot && S.m_dSuffix[e]>=iNeed )
					++e;
				return ( 4u<<10 ) | ( (uint32_t)e<<16 );
			};
			auto fnSetThr = [&] ( uint64_t uHiWord, uint64_t uLoWord )
			{
				tThrCur.m_uHi = uHiWord;
				tThrCur.m_uLo = uLoWord;
				const uint32_t uThrWx = (uint32_t)( uHiWord>>32 );
				// the threshold in the bound's fixed point: ( bound>>6 )*index weight >= K-th best weight
				int iThrFx = -( 1<<29 );	// no threshold yet / no usable bound: every present row is a candidate
				if ( S.m_bBound && uThrWx )
				{
					const int iW = q.m_iIndexWeight;
					const long long iThr = (long long)(int)( uThrWx ^ 0x80000000u );
					long long iQ = iThr>=0 ? ( iThr+iW-1 )/iW : -( ( -iThr )/iW );	// ceil ( thr / index weight )
					iQ = iQ<-( 1<<22 ) ? -( 1<<22 ) : iQ>( 1<<24 ) ? ( 1<<24 ) : iQ;
					iThrFx = (int)( iQ*64 );
				}
				__syncwarp();
				for ( int iTry=0; iTry<2; ++iTry )
				{
					// second try (too many combinations): the levels are ignored, one class per rank class
					const int nL1 = ( nNeg>=1 && !iTry ) ? 4 : 1, nL2 = ( nNeg>=2 && !iTry ) ? 4 : 1;
					const int nCombos = ( nFv-1 )*nL1*nL2;
					int n = 0;
					for ( int iBase=0; iBase<nCombos; iBase+=32 )
					{
						const int c = iBase+iLane;
						uint32_t uEntry = 0;
						if ( c<nCombos )
						{
							const int fv = 1 + c/( nL1*nL2 ), l1 = ( c/nL2 ) % nL1, l2 = c % nL2;
							const int iBaseNeed = iThrFx - S.m_dRankUb[fv];
							// even the rows owing the most need nothing: one class for the whole rank class
							const bool bAll = iBaseNeed + ( nL1>1 ? S.m_dNegPen[0][3] : 0 ) + ( nL2>1 ? S.m_dNegPen[1][3] : 0 )<=0;
							if ( bAll )
								uEntry = ( l1 | l2 ) ? 0u : ( (uint32_t)fv | ( 4u<<4 ) | ( 4u<<7 ) | ( 2u<<10 ) );
							else
							{
								const uint32_t uMode = fnMode ( iBaseNeed + ( nL1>1 ? S.m_dNegPen[0][l1] : 0 ) + ( nL2>1 ? S.m_dNegPen[1][l2] : 0 ) );
								if ( uMode )
									uEntry = (uint32_t)fv | ( (uint32_t)( nL1>1 ? l1 : 4 )<<4 ) | ( (uint32_t)( nL2>1 ? l2 : 4 )<<7 ) | uMode;
							}
						}
						const unsigned m = __ballot_sync ( FULL_MASK, uEntry!=0 );
						const int iSlot = n + __popc ( m & ( ( 1u<<iLane )-1u ) );
						if ( uEntry && iSlot<OB_MAX_CLASSES )
							pClass[iSlot] = uEntry;
						n += __popc ( m );
					}
					nClasses = n;
					if ( n<=OB_MAX_CLASSES )
						break;
				}
				__syncwarp();
			};
			fnSetThr ( tThr.m_uHi, tThr.m_uLo );

			while ( uMini<uMini1 )
			{
				if ( *( (volatile int *)&S.m_iPoolCnt )>=iTrigger )
					break;	// time to compact the pool: all warps meet at the barrier above
				if ( ( uMini & 3u )==3u )
				{
					// another item of this query may have raised the shared lower bound of the K-th best key meanwhile
					const unsigned long long uShared = *( (volatile unsigned long long *)( P.m_pQueryThr+tItem.m_uQuery ) );
					if ( uShared>tThrCur.m_uHi )
						fnSetThr ( uShared, 0 );
				}
				const uint32_t uLo = tItem.m_uRowLo + uMini*OB_MINI;
				const uint32_t uHi = min ( uLo+(uint32_t)OB_MINI, tItem.m_uRowHi );

				// no hot keyword: jump to the next mini-tile a sparse keyword touches
				if ( !nHot )
				{
					uint32_t uNext = 0xFFFFFFFFu;
					for ( int iSp=0; iSp<nSparseOps; ++iSp )
						uNext = min ( uNext, S.m_dNext[iWarp][q.m_dOps[S.m_dSparseOp[iSp]].m_uLeaf] );
					if ( uNext>=uHi )
					{
						if ( uNext==0xFFFFFFFFu || uNext>=tItem.m_uRowHi )
							uMini = uMini1;
						else
							uMini = max ( uMini+1, ( uNext-tItem.m_uRowLo )/OB_MINI );
						continue;
					}
				}

				// sparse keywords' postings of this mini-tile, gathered into a short per-warp list in op order
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
						S.m_dOpStart[iWarp][nSparseOps] = (uint16_t)nList;
					__syncwarp();
					// the next mini-tile any sparse keyword can touch
					uNextSparse = 0xFFFFFFFFu;
					for ( int iSp=0; iSp<nSparseOps; ++iSp )
						uNextSparse = min ( uNextSparse, S.m_dNext[iWarp][q.m_dOps[S.m_dSparseOp[iSp]].m_uLeaf] );
				}

				// the bitmap lines two mini-tiles ahead (one 128 B line per keyword and field)
				const uint32_t uWord = ( uLo>>5 ) + iLane;
				if ( uLo+3*OB_MINI<=tItem.m_uRowHi )
					for ( int i=iLane; i<nHot*4; i+=32 )
						if ( ( S.m_dBitFields[i>>2]>>( i & 3 ) ) & 1u )
							asm volatile ( "prefetch.global.L2 [%0];" :: "l" ( S.m_dBitPtr[i>>2] + ( i & 3 )*iBitStride + ( uLo>>5 ) + 64 ) );

				// sparse overlay: per-field bitmaps of the listed postings, and the rows holding any
				uint32_t dF[4] = { 0, 0, 0, 0 };
				uint32_t uSparseRows = 0;
				if ( nList )
				{
					#pragma unroll
					for ( int k=0; k<5; ++k )
						dOv[k][iLane] = 0;
					__syncwarp();
					for ( int e=iLane; e<nList; e+=32 )
					{
						const uint32_t uSlot = pList[e].m_uRowid, uFields = pList[e].m_uFields;
						const uint32_t uBit = 1u<<( uSlot & 31u );
						atomicOr ( &dOv[4][uSlot>>5], uBit );
						#pragma unroll
						for ( int f=0; f<4; ++f )
							if ( ( uFields>>f ) & 1u )
								atomicOr ( &dOv[f][uSlot>>5], uBit );
					}
					__syncwarp();
					#pragma unroll
					for ( int f=0; f<4; ++f )
						dF[f] = dOv[f][iLane];
					uSparseRows = dOv[4][iLane];
				}

				// 1. bitmap pass: four keywords at a time, all their loads in flight before the first use
				for ( int i0=0; i0<nHot; i0+=4 )
				{
					uint32_t dW[4][4];
					#pragma unroll
					for ( int j=0; j<4; ++j )
					{
						const bool b = i0+j<nHot;
						const uint32_t uMask = b ? S.m_dBitFields[i0+j] : 0u;
						const uint32_t * p = S.m_dBitPtr[b ? i0+j : 0] + uWord;
						#pragma unroll
						for ( int f=0; f<4; ++f )
							dW[j][f] = ( ( uMask>>f ) & 1u ) ? __ldg ( p + f*iBitStride ) : 0u;
					}
					#pragma unroll
					for ( int j=0; j<4; ++j )
						if ( i0+j<nHot )
						{
							#pragma unroll
							for ( int f=0; f<4; ++f )
								dF[f] |= dW[j][f];
							pPsm[( i0+j )*32+iLane] = dW[j][0] | dW[j][1] | dW[j][2] | dW[j][3];
						}
				}
				++uDbgMinis;
				const uint32_t uPresent = dF[0] | dF[1] | dF[2] | dF[3];
				iMyTotal += __popc ( uPresent );	// rows at/after the item's end are never present: store and bitmaps are zero there

				// 2. candidate rows per reachable class
				uint32_t uCand = 0;
				if ( __any_sync ( FULL_MASK, uPresent!=0 ) )
				{
					// tf levels of the negative keywords on this lane's rows: absent / 1 hit / 2-3 hits / >= 4 hits
					uint32_t dL1[4] = { 0, 0, 0, 0 }, dL2[4] = { 0, 0, 0, 0 };
					if ( nNeg>=1 )
					{
						const uint32_t uP = pPsm[S.m_dNegPsm[0]*32+iLane], u2 = __ldg ( S.m_dNegLvl[0]+uWord ), u4 = __ldg ( S.m_dNegLvl[0]+iBitStride+uWord );
						dL1[0] = ~uP; dL1[1] = uP & ~u2; dL1[2] = uP & u2 & ~u4; dL1[3] = uP & u4;
					}
					if ( nNeg>=2 )
					{
						const uint32_t uP = pPsm[S.m_dNegPsm[1]*32+iLane], u2 = __ldg ( S.m_dNegLvl[1]+uWord ), u4 = __ldg ( S.m_dNegLvl[1]+iBitStride+uWord );
						dL2[0] = ~uP; dL2[1] = uP & ~u2; dL2[2] = uP & u2 & ~u4; dL2[3] = uP & u4;
					}
					int iPrevFv = 0;
					uint32_t uFvRows = 0;
					for ( int iCls=0; iCls<nClasses; ++iCls )
					{
						const uint32_t uCode = pClass[iCls];	// (same address in every lane)
						const int fv = (int)( uCode & 15u ), l1 = (int)( ( uCode>>4 ) & 7u ), l2 = (int)( ( uCode>>7 ) & 7u );
						const uint32_t uSelMode = ( uCode>>10 ) & 7u;
						if ( fv!=iPrevFv )
						{
							uFvRows = uPresent;
							#pragma unroll
							for ( int f=0; f<4; ++f )
								uFvRows &= ( ( fv>>f ) & 1 ) ? dF[f] : ~dF[f];
							iPrevFv = fv;
						}
						uint32_t m = uFvRows;
						m &= l1==0 ? dL1[0] : l1==1 ? dL1[1] : l1==2 ? dL1[2] : l1==3 ? dL1[3] : 0xFFFFFFFFu;
						m &= l2==0 ? dL2[0] : l2==1 ? dL2[1] : l2==2 ? dL2[2] : l2==3 ? dL2[3] : 0xFFFFFFFFu;
						if ( uSelMode==1 )
							m &= uSparseRows;
						if ( !__any_sync ( FULL_MASK, m!=0 ) )
							continue;
						if ( uSelMode>=3 )
						{
							const int n = (int)( ( uCode>>16 ) & 255u );
							uint32_t uSel = pPsm[iLane];
							if ( uSelMode==3 )
								for ( int i=1; i<n; ++i )
									uSel &= pPsm[i*32+iLane];
							else
								for ( int i=1; i<n; ++i )
									uSel |= pPsm[i*32+iLane];
							m &= uSel | uSparseRows;
						}
						uCand |= m;
					}
				}

				if ( __any_sync ( FULL_MASK, uCand!=0 ) )
				{
					// 3a. candidate rows holding a sparse posting: compacted into the queue's free tail and evaluated now, all ops in order
					const uint32_t uNow = uCand & uSparseRows;
					uDbgSparse += __popc ( uNow );
					uDbgHot += __popc ( uCand & ~uSparseRows );
					if ( __any_sync ( FULL_MASK, uNow!=0 ) )
					{
						uint32_t * pCand = pQueue + 32;
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
							pCand[iOff++] = (uint32_t)( iLane*32 + __ffs ( m )-1 );
						__syncwarp();
						for ( int iBase=0; iBase<nCand; iBase+=32 )
						{
							const bool bAct = iBase+iLane<nCand;
							const int sRow = bAct ? (int)pCand[iBase+iLane] : 0;	// slot inside the mini-tile
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
						__syncwarp();
					}

					// 3b. the other candidate rows join the warp's queue (index-local rowids); full passes of 32 rows run now
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
							pQueue[iOff++] = uLo + iLane*32 + __ffs ( m )-1;
						nQueue += nNew;
						__syncwarp();
						while ( nQueue>=32 )
						{
							nQueue -= 32;
							fnExactHot ( true, pQueue[nQueue+iLane], pPool, tThrCur );
						}
						__syncwarp();
					}
				}
				++uMini;
			}
			// the queued candidate rows are evaluated against this round's threshold and pool buffer
			if ( nQueue )
			{
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
			if ( P.m_pDebug )
			{
				#pragma unroll
				for ( int d=16; d; d>>=1 )
				{
					uDbgHot += __shfl_xor_sync ( FULL_MASK, uDbgHot, d );
					uDbgSparse += __shfl_xor_sync ( FULL_MASK, uDbgSparse, d );
				}
				if ( iLane==0 )
				{
					atomicAdd ( P.m_pDebug+0, (unsigned long long)uDbgMinis );
					atomicAdd ( P.m_pDebug+1, (unsigned long long)uDbgHot );
					atomicAdd ( P.m_pDebug+2, (unsigned long long)uDbgSparse );
					atomicAdd ( P.m_pDebug+4, (unsigned long long)iMyTotal );
				}
			}
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
