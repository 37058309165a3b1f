// K3c orbits_kernel: launch class 5, OR programs under BM25 in relevance order ("OR over bit planes"): pure ORs of keywords, and
// ORs of AND groups whose keywords all sit in the hot store (a pure AND of hot keywords is one such group).
//
// A top-K OR query has to COUNT every matching row (total_found = every push, src/sphinxsort.cpp:724) but only has to RANK the
// rows that can still enter the top K. The batch's hot-term store carries, next to the u16 {hits, fields, tf class} per row, one
// presence BITMAP per (hot keyword, field): 1 bit per row instead of 16. A warp walks its share of the work item in 1024-row
// mini-tiles, one 32-bit word (32 rows) per lane and bitmap:
//
//  1. bitmap pass, ~1 instruction per 32 rows and keyword: per-field ORs give the matched-field mask F(row) bit-parallel, the
//     per-keyword presence words P_i go to shared memory; popc ( present ) is the exact total_found. Keywords outside the hot
//     store were decoded once per batch run into plain posting lists (sparse_decode_kernel): a per-warp cursor per keyword sets
//     the bits of the mini-tile's postings in a small shared-memory overlay (per field + "row holds a listed posting").
//  2. candidate selection, MaxScore style (Turtle & Flood), bit-parallel: a row's weight is at most
//     rank ( F )*1000 + 500 + 1000*sum over the PRESENT keywords of max ( idf_i, 0 ) (tf < 1; ExtRanker_WeightSum_c,
//     src/sphinxsearch.cpp:1096-1141). For every rank class fv (value of F) the warp knows how much TF*IDF a row of that class needs
//     to reach the K-th best weight so far; with the hot keywords sorted by their bound, the rows that can still make it are
//     those holding ALL keywords of a prefix ("required": without one of them the rest cannot reach it) or, if none is required,
//     ANY keyword of the "essential" prefix (the rest together stays below it). Rows holding a listed posting always
//     qualify while their class can reach the threshold at all.
//     Keywords with a NEGATIVE idf (df > N/2: the stop words) lower the weight of nearly every row by about as much as the
//     rare keywords raise it, so ignoring them leaves ~20 % of the rows as candidates (scripts/or_bound_study.py). Up to two of
//     them refine the classes by their tf level on the row (absent / 1 hit / 2-3 hits / >= 4 hits, from two more bitmaps the
//     store keeps for keywords in >= 1/3 of the rows): a row of level l owes at least tf_min ( l )*|idf|, which raises the
//     TF*IDF its positive keywords must bring. Selections only get stricter as the levels go up, so a class is listed as
//     "rows of levels <= (a, b)" and dropped when a neighbour (a+1, b) / (a, b+1) selects the same way; the list is rebuilt
//     whenever the warp's threshold moves. ~1 % of the rows stay candidates.
//  3. exact pass for the candidate rows only, 32 at a time, one lane per row, TF*IDF in op order (ExtOr_c left fold,
//     src/searchnode.cpp:3486-3504): hot keywords from the u16 store, listed ones by a binary search of the part of the posting
//     list the warp has walked; ranked and pushed as in K3.
//
// Everything around it (work items, candidate pool, on-demand CTA radix select, the query's shared K-th-best bound) is that
// of stream_kernel<512,1>, which stays available for A/B runs (mgpu_index_set_option "or_bits" = 0).
#pragma once
// (included from kernels.cu inside namespace mgpu, after stream_kernel.cuh)

static const int OB_MINI = 1024;					///< rows per warp step: one bitmap word per lane
static const int OB_QUEUE = 32 + OB_MINI;			///< per-warp queue of candidate rows
static const size_t OB_WARP_SMEM = ( MAX_LEAVES*32 + OB_QUEUE )*4;
static const int OB_MAX_CLASSES = 64;				///< (rank class, level, level) combinations listed per warp; more -> levels are ignored
static const uint32_t OB_LISTED = 0x80000000u;		///< queue entry: the row holds a listed (non-hot) posting

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
	// bitmap pass: the program's hot UNITS (a hot keyword of the OR, or an AND group of hot keywords) sorted by their weight bound,
	// biggest first; their member keywords flat in that order
	const uint32_t * m_dBitPtr[MAX_LEAVES];				///< member: the keyword's field-0 bitmap (field f: + f*bit stride)
	uint32_t		m_dBitFields[MAX_LEAVES];			///< member: queried fields the index has | 256 if it is the last one of its unit; padded with 0 to a multiple of 4 members
	int32_t			m_nMembers;
	int32_t			m_dUb[MAX_LEAVES];					///< unit: sum over its keywords of ceil ( max ( idf, 0 )*64000 ) + 1
	int32_t			m_dSuffix[MAX_LEAVES+1];			///< sum of m_dUb[i..]
	uint8_t			m_dSortLeaf[MAX_LEAVES];			///< unit: its leaf when it is a single keyword, 0xFF for an AND group
	int32_t			m_nHot;								///< hot units
	int32_t			m_iUbListed;						///< the same bound summed over the listed keywords
	// penalty classes: up to two hot keywords with idf < 0 that own tf-level bitmaps
	int32_t			m_nNeg;
	int32_t			m_dNegPsm[2];						///< the keyword's entry in the sorted list (its presence word)
	const uint32_t * m_dNegLvl[2];						///< its ">= 2 hits" bitmap (">= 4 hits": + bit stride)
	int32_t			m_dNegPen[2][4];					///< what a row of tf level l owes at least, in the bound's fixed point
	uint32_t		m_dClass[EVAL_WARPS][OB_MAX_CLASSES];	///< fv | a<<4 | b<<7 | mode<<10 | prefix length<<16: rows of rank class fv and levels <= (a, b)
	// exact pass: op order
	const uint16_t * m_dOpPtr[MAX_LEAVES];				///< the op's row of the u16 store (null = listed keyword)
	uint8_t			m_dOpList[MAX_LEAVES];				///< listed keyword: its entry in the arrays below
	uint8_t			m_dOpLast[MAX_LEAVES];				///< the op closes its unit
	uint8_t			m_dHotLeaf[MAX_LEAVES];				///< hot keywords only, op order
	uint8_t			m_dHotLast[MAX_LEAVES];				///< ... closes its unit
	const uint16_t * m_dHotPtr[MAX_LEAVES+4];
	int32_t			m_nHotOps;
	// listed keywords
	uint8_t			m_dListLeaf[MAX_LEAVES];
	uint32_t		m_dListEnd[MAX_LEAVES];				///< one past the keyword's last entry
	int32_t			m_nListed;
	uint32_t		m_dListBeg[EVAL_WARPS][MAX_LEAVES];	///< first entry at/after the warp's first row
	uint32_t		m_dListCur[EVAL_WARPS][MAX_LEAVES];	///< first entry at/after the warp's position
	uint32_t		m_dListNext[EVAL_WARPS][MAX_LEAVES];	///< its rowid (0xFFFFFFFF = exhausted)
	float			m_dTf[256];
	uint32_t		m_dOv[EVAL_WARPS][5][32];			///< listed postings of the mini-tile: per-field bitmaps [0..3], rows holding any [4]
};

/// first entry of pRows[uFrom, uTo) with a rowid >= uRow (uTo if none); every lane gets the result
__device__ __forceinline__ uint32_t ListLowerBound ( const uint32_t * __restrict__ pRows, uint32_t uFrom, uint32_t uTo, uint32_t uRow, int iLane )
{
	uint32_t a = uFrom, b = uTo;
	while ( b-a>32u )
	{
		// 32 probes split [a, b) into 33 parts
		const uint32_t uStep = ( b-a )/33u + 1u;
		const uint32_t i = a + ( iLane+1 )*uStep;
		const bool bGe = i>=b || __ldg ( pRows+i )>=uRow;
		const unsigned m = __ballot_sync ( FULL_MASK, bGe );
		const int k = m ? __ffs ( m )-1 : 32;	// probes 0..k-1 are below uRow, probe k is not
		const uint32_t uNewA = k ? a + k*uStep + 1u : a;
		b = k<32 ? min ( b, a + ( k+1 )*uStep ) : b;
		a = min ( uNewA, b );
	}
	const uint32_t i = a+iLane;
	const bool bGe = i>=b || __ldg ( pRows+i )>=uRow;
	const unsigned m = __ballot_sync ( FULL_MASK, bGe );
	return min ( b, a + ( m ? __ffs ( m )-1 : 32 ) );
}

/// one lane: value of rowid uRow in pRows[uFrom, uTo) (sorted), 0 if absent
__device__ __forceinline__ uint32_t ListFind ( const DevPostingLists_t & tLists, uint32_t uFrom, uint32_t uTo, uint32_t uRow )
{
	uint32_t lo = uFrom, hi = uTo;
	while ( lo<hi )
	{
		const uint32_t mid = lo + ( ( hi-lo )>>1 );
		if ( __ldg ( tLists.m_pRows+mid )<uRow ) lo = mid+1; else hi = mid;
	}
	return ( lo<uTo && __ldg ( tLists.m_pRows+lo )==uRow ) ? __ldg ( tLists.m_pVals+lo ) : 0u;
}

/// NF = 2: indexes with <= 2 fields, 4: three or four
template<int NF>
__global__ void __launch_bounds__ ( EVAL_THREADS, 3 ) orbits_kernel ( EvalParams_t P )
{
	extern __shared__ __align__(16) uint8_t dDyn[];
	__shared__ OrBitsShared_t S;
	const int tid = threadIdx.x, iWarp = tid>>5, iLane = tid & 31;
	const DevIndex_t & tIdx = P.m_tIndex;
	const DevPostingLists_t & tLists = P.m_tLists;

	uint32_t * pPsm = reinterpret_cast<uint32_t *>( dDyn + (size_t)iWarp*OB_WARP_SMEM );	// [MAX_LEAVES][32] presence words of the hot keywords
	uint32_t * pQueue = pPsm + MAX_LEAVES*32;												// [OB_QUEUE]
	Key128_t * pPool0 = P.m_pPool + (size_t)blockIdx.x*2*P.m_iPoolCap;
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
			LoadQuery ( S.m_tQ, P.m_pQueries, P.m_pQueryExt, tItem.m_uQuery, tid, EVAL_THREADS );
		}
		if ( tid==0 )
		{
			S.m_iPoolCnt = 0;
			S.m_iPoolBuf = 0;
			S.m_tThr.m_uHi = 0; S.m_tThr.m_uLo = 0;
			S.m_uTotal = 0;
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
			int nHot = 0, nHotOps = 0, nListed = 0, iUbListed = 0;
			const uint32_t uIndexFields = ( 1u<<P.m_tHot.m_nBitFields )-1u;
			// units of the OR fold, in program order: AND groups (hot keywords only: the host's test) or single keywords
			const bool bGroups = q.m_nGroups>0;
			const int nUnits = bGroups ? q.m_nGroups : q.m_nOps;
			uint8_t dUnitOp0[MAX_LEAVES], dUnitOps[MAX_LEAVES], dOrder[MAX_LEAVES];
			int dUnitUb[MAX_LEAVES];
			for ( int u=0; u<nUnits && u<MAX_LEAVES; ++u )
			{
				const int iOp0 = bGroups ? q.m_dGroupOp0[u] : u, nUnitOps = bGroups ? q.m_dGroupOps[u] : 1;
				int iUbUnit = 0;
				bool bListed = false;
				for ( int iOp=iOp0; iOp<iOp0+nUnitOps; ++iOp )
				{
					const int l = q.m_dOps[iOp].m_uLeaf;
					const DevLeaf_t & tLeaf = q.m_dLeaves[l];
					const float fIDF = tLeaf.m_fIDF;
					// a negated keyword at the end of an AND group (`a b -c`): rows that hold it leave the unit, it adds no weight (flag 2)
					const bool bNegOp = q.m_dOps[iOp].m_eCode==OP_TERM_ANDNOT;
					// tf < 1: the keyword adds less than idf when it sits on the row, and nothing above 0 when idf <= 0
					iUbUnit += ( fIDF>0.0f && !bNegOp ) ? (int)ceilf ( __fmul_rn ( fminf ( fIDF, 1.0f ), 64000.0f ) )+1 : 0;
					S.m_dOpPtr[iOp] = nullptr;
					S.m_dOpLast[iOp] = (uint8_t)( ( ( iOp==iOp0+nUnitOps-1 ) ? 1 : 0 ) | ( bNegOp ? 2 : 0 ) );
					if ( tLeaf.m_iHot<0 )
					{
						// (a single keyword outside the hot store: its postings come from the decoded list)
						bListed = true;
						S.m_dOpList[iOp] = (uint8_t)nListed;
						S.m_dListLeaf[nListed] = (uint8_t)l;
						S.m_dListEnd[nListed] = tLeaf.m_nBlocks ? tLeaf.m_uListOff+tLeaf.m_nDocs : 0u;	// (a keyword the index does not hold has no list)
						++nListed;
						continue;
					}
					const uint16_t * pRow = P.m_tHot.m_pData + (size_t)tLeaf.m_iHot*P.m_tHot.m_iStride;
					S.m_dOpPtr[iOp] = pRow;
					S.m_dHotPtr[nHotOps] = pRow;
					S.m_dHotLast[nHotOps] = S.m_dOpLast[iOp];
					S.m_dHotLeaf[nHotOps++] = (uint8_t)l;
				}
				if ( bListed )
				{
					iUbListed += iUbUnit;
					continue;
				}
				// insertion into the order sorted by bound, biggest first
				dUnitOp0[nHot] = (uint8_t)iOp0; dUnitOps[nHot] = (uint8_t)nUnitOps; dUnitUb[nHot] = iUbUnit;
				int j = nHot++;
				for ( ; j>0 && dUnitUb[dOrder[j-1]]<iUbUnit; --j )
					dOrder[j] = dOrder[j-1];
				dOrder[j] = (uint8_t)( nHot-1 );
			}
			int nMem = 0;
			for ( int i=0; i<nHot; ++i )
			{
				const int u = dOrder[i];
				S.m_dUb[i] = dUnitUb[u];
				S.m_dSortLeaf[i] = dUnitOps[u]==1 ? q.m_dOps[dUnitOp0[u]].m_uLeaf : (uint8_t)0xFF;
				for ( int iOp=dUnitOp0[u]; iOp<dUnitOp0[u]+dUnitOps[u]; ++iOp )
				{
					const DevLeaf_t & tLeaf = q.m_dLeaves[q.m_dOps[iOp].m_uLeaf];
					S.m_dBitPtr[nMem] = P.m_tHot.m_pBits + (size_t)tLeaf.m_iHot*P.m_tHot.m_nBitFields*iBitStride;
					S.m_dBitFields[nMem] = ( tLeaf.m_uQueriedFields & uIndexFields & 255u ) | ( ( iOp==dUnitOp0[u]+dUnitOps[u]-1 ) ? 256u : 0u )
						| ( q.m_dOps[iOp].m_eCode==OP_TERM_ANDNOT ? 512u : 0u );
					++nMem;
				}
			}
			S.m_nMembers = nMem;
			for ( int i=nMem; i<( ( nMem+3 ) & ~3 ); ++i )
			{
				// (padding: no field, no load; the last unit has been closed before)
				S.m_dBitPtr[i] = P.m_tHot.m_pBits;
				S.m_dBitFields[i] = 0;
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
				if ( S.m_dSortLeaf[i]==0xFF )
					continue;	// (a stop word inside an AND group only lowers that group's sum: ignored by the bound)
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
			S.m_nListed = nListed;
			S.m_iUbListed = iUbListed;
		}
		int iMyTotal = 0;
		uint32_t uDbgMinis = 0, uDbgHot = 0, uDbgListed = 0;

		// this warp's contiguous share of the item
		const uint32_t nMinis = ( tItem.m_uRowHi-tItem.m_uRowLo+OB_MINI-1 )/OB_MINI;
		const uint32_t uMini0 = (uint32_t)( (uint64_t)nMinis*iWarp/EVAL_WARPS ), uMini1 = (uint32_t)( (uint64_t)nMinis*( iWarp+1 )/EVAL_WARPS );
		uint32_t uMini = uMini0;
		const bool bAnyEscape = P.m_tHot.m_pEscapeCount && __ldg ( P.m_tHot.m_pEscapeCount )!=0;

		__syncthreads();	// S.m_nHot and friends
		int nQueue = 0;
		const int nOps = q.m_nOps, nHot = S.m_nHot, nHotOps = S.m_nHotOps, nListed = S.m_nListed, nMembers = S.m_nMembers;
		const int iUbHot = S.m_dSuffix[0], iUbListed = S.m_iUbListed;
		const int iIndexWeight = q.m_iIndexWeight;
		const int nFv = 1<<P.m_tHot.m_nBitFields;
		const int nNeg = S.m_nNeg;
		uint32_t * pClass = S.m_dClass[iWarp];

		// listed keywords: where this warp's rows start in each posting list
		uint32_t uNextListed = 0xFFFFFFFFu;
		for ( int iSp=0; iSp<nListed; ++iSp )
		{
			const DevLeaf_t & tLeaf = q.m_dLeaves[S.m_dListLeaf[iSp]];
			const uint32_t uEnd = S.m_dListEnd[iSp];
			uint32_t uBeg = uEnd, uNext = 0xFFFFFFFFu;
			if ( uEnd && uMini0<uMini1 )
			{
				uBeg = ListLowerBound ( tLists.m_pRows, tLeaf.m_uListOff, uEnd, tItem.m_uRowLo + uMini0*OB_MINI, iLane );
				if ( uBeg<uEnd )
					uNext = __ldg ( tLists.m_pRows+uBeg );
			}
			if ( iLane==0 )
			{
				S.m_dListBeg[iWarp][iSp] = uBeg;
				S.m_dListCur[iWarp][iSp] = uBeg;
				S.m_dListNext[iWarp][iSp] = uNext;
			}
			uNextListed = min ( uNextListed, uNext );
		}
		__syncwarp();

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
		// exact TF*IDF of one queued row per lane, in op order. Rows without a listed posting (the bulk) only read the hot keywords.
		auto fnExact = [&] ( bool bAct, uint32_t uEntry, Key128_t * pPool, const Key128_t & tThr )
		{
			const uint32_t uRow = uEntry & ~OB_LISTED;
			float fT = 0.0f;
			uint32_t uF = 0;
			bool bPres = false;
			if ( !__any_sync ( FULL_MASK, bAct && ( uEntry & OB_LISTED ) ) )
			{
				// unit by unit in program order: an AND group counts when all its keywords sit on the row, its sum folded keyword by
				// keyword (ExtAnd_c / ExtMultiAnd_T: left + right); the units' sums are folded by ExtOr_c (src/searchnode.cpp:3486-3504)
				float fG = 0.0f;
				uint32_t uFG = 0;
				bool bAll = true, bFirst = true;
				for ( int h0=0; h0<nHotOps; h0+=4 )
				{
					uint32_t dRaw[4];
					#pragma unroll
					for ( int i=0; i<4; ++i )
						dRaw[i] = ( bAct && h0+i<nHotOps ) ? __ldg ( S.m_dHotPtr[h0+i]+uRow ) : 0u;
					#pragma unroll
					for ( int i=0; i<4; ++i )
					{
						if ( h0+i>=nHotOps )
							break;
						const uint32_t uHits = dRaw[i] & 255u;
						const DevLeaf_t & tLeaf = q.m_dLeaves[S.m_dHotLeaf[h0+i]];
						const uint32_t uFields = ( dRaw[i]>>8 ) & tLeaf.m_uQueriedFields;
						if ( S.m_dHotLast[h0+i] & 2 )
						{
							if ( uHits && uFields )
								bAll = false;	// the row holds a negated keyword of the group
						} else if ( !uHits || !uFields )
							bAll = false;
						else
						{
							float fBase = S.m_dTf[uHits];
							if ( bAnyEscape && uHits==255 )
								fBase = HotEscapeTf ( P.m_tHot, tLeaf.m_iHot, uRow );
							const float fTf = __fmul_rn ( fBase, tLeaf.m_fIDF );
							fG = bFirst ? fTf : __fadd_rn ( fG, fTf );
							uFG |= uFields;
							bFirst = false;
						}
						if ( S.m_dHotLast[h0+i] & 1 )
						{
							if ( bAll && !bFirst )
							{
								fT = bPres ? __fadd_rn ( fT, fG ) : fG;
								uF |= uFG;
								bPres = true;
							}
							fG = 0.0f; uFG = 0; bAll = true; bFirst = true;
						}
					}
				}
			} else
			{
				const bool bListed = bAct && ( uEntry & OB_LISTED );
				float fG = 0.0f;
				uint32_t uFG = 0;
				bool bAll = true, bFirst = true;
				const bool bGroups = q.m_nGroups>0;
				const int nUnits = bGroups ? q.m_nGroups : nOps;
				for ( int u=0; u<nUnits; ++u )
				for ( int iOp=( bGroups ? q.m_dGroupOp0[u] : u ), iOpEnd=iOp+( bGroups ? q.m_dGroupOps[u] : 1 ); iOp<iOpEnd; ++iOp )
				{
					const DevLeaf_t & tLeaf = q.m_dLeaves[q.m_dOps[iOp].m_uLeaf];
					const uint16_t * pRow = S.m_dOpPtr[iOp];
					uint32_t uHits = 0, uFields = 0;
					float fBase = 0.0f;
					if ( pRow )
					{
						const uint32_t uRaw = bAct ? __ldg ( pRow+uRow ) : 0u;
						uHits = uRaw & 255u;
						uFields = ( uRaw>>8 ) & tLeaf.m_uQueriedFields;
						fBase = S.m_dTf[uHits];
						if ( bAnyEscape && uHits==255 )
							fBase = HotEscapeTf ( P.m_tHot, tLeaf.m_iHot, uRow );
					} else if ( bListed )
					{
						// the row lies in the part of the list this warp has walked
						const int iSp = S.m_dOpList[iOp];
						const uint32_t uVal = ListFind ( tLists, S.m_dListBeg[iWarp][iSp], S.m_dListCur[iWarp][iSp], uRow );
						uHits = uVal & 0xFFFFFFu;
						uFields = ( uVal>>24 ) & tLeaf.m_uQueriedFields;
						// ExtTerm_T::GetDocsChunk, src/searchnode.cpp:1946
						const float fHits = __uint2float_rn ( uHits );
						fBase = uHits<255u ? S.m_dTf[uHits] : __fdiv_rn ( fHits, __fadd_rn ( fHits, 1.2f ) );
					}
					if ( S.m_dOpLast[iOp] & 2 )
					{
						if ( uHits && uFields )
							bAll = false;
					} else if ( !uHits || !uFields )
						bAll = false;
					else
					{
						const float fTf = __fmul_rn ( fBase, tLeaf.m_fIDF );
						fG = bFirst ? fTf : __fadd_rn ( fG, fTf );
						uFG |= uFields;
						bFirst = false;
					}
					if ( S.m_dOpLast[iOp] & 1 )
					{
						if ( bAll && !bFirst )
						{
							fT = bPres ? __fadd_rn ( fT, fG ) : fG;
							uF |= uFG;
							bPres = true;
						}
						fG = 0.0f; uFG = 0; bAll = true; bFirst = true;
					}
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
			int nClasses = 0;	// listed (rank class, level bound of negative keyword 1, of negative keyword 2) combinations in pClass
			// how the rows of a class are selected given the TF*IDF its positive keywords must bring: mode 0 = no row can reach the
			// threshold, 1 = only rows holding a listed posting, 2 = every row, 3 = rows holding all of the first n hot keywords,
			// 4 = rows holding any of the first n hot keywords (3, 4: or a listed posting). Stricter as the need grows.
			auto fnMode = [&] ( int iNeed ) -> uint32_t
			{
				if ( iNeed<=0 )
					return 2u<<10;
				if ( iNeed>iUbHot+iUbListed )
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
							const int fv = 1 + c/( nL1*nL2 ), a = ( c/nL2 ) % nL1, b = c % nL2;
							const int iBaseNeed = iThrFx - S.m_dRankUb[fv];
							const int iPenA = nL1>1 ? S.m_dNegPen[0][a] : 0, iPenB = nL2>1 ? S.m_dNegPen[1][b] : 0;
							const uint32_t uMode = fnMode ( iBaseNeed + iPenA + iPenB );
							// rows of levels <= (a, b) under this selection; the same selection one level up covers them
							bool bKeep = uMode!=0;
							if ( bKeep && a+1<nL1 )
								bKeep = fnMode ( iBaseNeed + S.m_dNegPen[0][a+1] + iPenB )!=uMode;
							if ( bKeep && b+1<nL2 )
								bKeep = fnMode ( iBaseNeed + iPenA + S.m_dNegPen[1][b+1] )!=uMode;
							if ( bKeep )
								uEntry = (uint32_t)fv | ( (uint32_t)( nL1>1 ? a : 3 )<<4 ) | ( (uint32_t)( nL2>1 ? b : 3 )<<7 ) | uMode;
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

				// no hot keyword: jump to the next mini-tile a listed keyword touches
				if ( !nHot && uNextListed>=uHi )
				{
					if ( uNextListed==0xFFFFFFFFu || uNextListed>=tItem.m_uRowHi )
						uMini = uMini1;
					else
						uMini = max ( uMini+1, ( uNextListed-tItem.m_uRowLo )/OB_MINI );
					continue;
				}

				// the bitmap lines two mini-tiles ahead (one 128 B line per keyword and field)
				const uint32_t uWord = ( uLo>>5 ) + iLane;
				if ( uLo+3*OB_MINI<=tItem.m_uRowHi )
					for ( int i=iLane; i<nMembers*NF; i+=32 )
						if ( ( S.m_dBitFields[i/NF]>>( i%NF ) ) & 1u )
							asm volatile ( "prefetch.global.L2 [%0];" :: "l" ( S.m_dBitPtr[i/NF] + ( i%NF )*iBitStride + ( uLo>>5 ) + 64 ) );

				// listed keywords' postings of this mini-tile -> overlay: per-field bitmaps, and the rows holding any
				uint32_t dF[NF];
				#pragma unroll
				for ( int f=0; f<NF; ++f )
					dF[f] = 0;
				uint32_t uListedRows = 0;
				if ( uNextListed<uHi )
				{
					#pragma unroll
					for ( int k=0; k<5; ++k )
						dOv[k][iLane] = 0;
					__syncwarp();
					uNextListed = 0xFFFFFFFFu;
					for ( int iSp=0; iSp<nListed; ++iSp )
					{
						uint32_t uNext = S.m_dListNext[iWarp][iSp];
						if ( uNext<uHi )
						{
							const uint32_t uQueried = q.m_dLeaves[S.m_dListLeaf[iSp]].m_uQueriedFields;
							const uint32_t uEnd = S.m_dListEnd[iSp];
							uint32_t uCur = S.m_dListCur[iWarp][iSp];
							while ( true )
							{
								const uint32_t i = uCur+iLane;
								const uint32_t r = i<uEnd ? __ldg ( tLists.m_pRows+i ) : 0xFFFFFFFFu;
								const bool bIn = r<uHi;
								if ( bIn )
								{
									const uint32_t uFields = ( __ldg ( tLists.m_pVals+i )>>24 ) & uQueried;
									if ( uFields )
									{
										const uint32_t uSlot = r-uLo, uBit = 1u<<( uSlot & 31u );
										atomicOr ( &dOv[4][uSlot>>5], uBit );
										#pragma unroll
										for ( int f=0; f<NF; ++f )
											if ( ( uFields>>f ) & 1u )
												atomicOr ( &dOv[f][uSlot>>5], uBit );
									}
								}
								const int n = __popc ( __ballot_sync ( FULL_MASK, bIn ) );
								uCur += n;
								if ( n<32 )
								{
									uNext = __shfl_sync ( FULL_MASK, r, n );
									break;
								}
							}
							__syncwarp();
							if ( iLane==0 )
							{
								S.m_dListCur[iWarp][iSp] = uCur;
								S.m_dListNext[iWarp][iSp] = uNext;
							}
						}
						uNextListed = min ( uNextListed, uNext );
					}
					__syncwarp();
					#pragma unroll
					for ( int f=0; f<NF; ++f )
						dF[f] = dOv[f][iLane];
					uListedRows = dOv[4][iLane];
				}

				// 1. bitmap pass: four keywords at a time, all their loads in flight before the first use. A unit's presence word is the AND of
				// its keywords' presence words; its fields only count where the whole unit sits on the row.
				{
					uint32_t uUnitP = 0xFFFFFFFFu, dFu[NF];
					#pragma unroll
					for ( int f=0; f<NF; ++f )
						dFu[f] = 0;
					int iUnit = 0;
					for ( int i0=0; i0<nMembers; i0+=4 )
					{
						uint32_t dW[4][NF], dInfo[4];
						#pragma unroll
						for ( int j=0; j<4; ++j )
						{
							dInfo[j] = S.m_dBitFields[i0+j];
							const uint32_t * p = S.m_dBitPtr[i0+j] + uWord;
							#pragma unroll
							for ( int f=0; f<NF; ++f )
								dW[j][f] = ( ( dInfo[j]>>f ) & 1u ) ? __ldg ( p + f*iBitStride ) : 0u;
						}
						#pragma unroll
						for ( int j=0; j<4; ++j )
						{
							uint32_t uAny = 0;
							#pragma unroll
							for ( int f=0; f<NF; ++f )
								uAny |= dW[j][f];
							if ( dInfo[j] & 512u )
								uUnitP &= ~uAny;	// negated member: its rows leave the unit
							else
							{
								#pragma unroll
								for ( int f=0; f<NF; ++f )
									dFu[f] |= dW[j][f];
								uUnitP &= uAny;
							}
							if ( dInfo[j] & 256u )
							{
								#pragma unroll
								for ( int f=0; f<NF; ++f )
								{
									dF[f] |= dFu[f] & uUnitP;
									dFu[f] = 0;
								}
								pPsm[iUnit*32+iLane] = uUnitP;
								++iUnit;
								uUnitP = 0xFFFFFFFFu;
							}
						}
					}
				}
				++uDbgMinis;
				uint32_t uPresent = 0;
				#pragma unroll
				for ( int f=0; f<NF; ++f )
					uPresent |= dF[f];
				iMyTotal += __popc ( uPresent );	// rows at/after the item's end are never present: store, bitmaps and lists hold none

				// 2. candidate rows per listed class
				uint32_t uCand = 0;
				if ( __any_sync ( FULL_MASK, uPresent!=0 ) )
				{
					// rows whose negative keywords sit at tf level <= 0 (absent), <= 1 (1 hit), <= 2 (2-3 hits), any
					uint32_t dT1[3] = { 0, 0, 0 }, dT2[3] = { 0, 0, 0 };
					if ( nNeg>=1 )
					{
						const uint32_t uP = pPsm[S.m_dNegPsm[0]*32+iLane], u2 = __ldg ( S.m_dNegLvl[0]+uWord ), u4 = __ldg ( S.m_dNegLvl[0]+iBitStride+uWord );
						dT1[0] = ~uP; dT1[1] = ~( uP & u2 ); dT1[2] = ~( uP & u4 );
					}
					if ( nNeg>=2 )
					{
						const uint32_t uP = pPsm[S.m_dNegPsm[1]*32+iLane], u2 = __ldg ( S.m_dNegLvl[1]+uWord ), u4 = __ldg ( S.m_dNegLvl[1]+iBitStride+uWord );
						dT2[0] = ~uP; dT2[1] = ~( uP & u2 ); dT2[2] = ~( uP & u4 );
					}
					int iPrevFv = 0;
					uint32_t uFvRows = 0;
					bool bFvAny = false;
					for ( int iCls=0; iCls<nClasses; ++iCls )
					{
						const uint32_t uCode = pClass[iCls];	// (same address in every lane)
						const int fv = (int)( uCode & 15u ), a = (int)( ( uCode>>4 ) & 7u ), b = (int)( ( uCode>>7 ) & 7u );
						const uint32_t uSelMode = ( uCode>>10 ) & 7u;
						if ( fv!=iPrevFv )
						{
							uFvRows = uPresent;
							#pragma unroll
							for ( int f=0; f<NF; ++f )
								uFvRows &= ( ( fv>>f ) & 1 ) ? dF[f] : ~dF[f];
							iPrevFv = fv;
							bFvAny = __any_sync ( FULL_MASK, uFvRows!=0 );
						}
						if ( !bFvAny )
							continue;	// no row of this mini-tile matches in exactly these fields
						uint32_t m = uFvRows;
						m &= a==0 ? dT1[0] : a==1 ? dT1[1] : a==2 ? dT1[2] : 0xFFFFFFFFu;
						m &= b==0 ? dT2[0] : b==1 ? dT2[1] : b==2 ? dT2[2] : 0xFFFFFFFFu;
						m &= ~uCand;
						if ( uSelMode==1 )
							m &= uListedRows;
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
							m &= uSel | uListedRows;
						}
						uCand |= m;
					}
				}

				// 3. the candidate rows join the warp's queue (index-local rowid, flag: holds a listed posting); full passes of 32 run now
				if ( __any_sync ( FULL_MASK, uCand!=0 ) )
				{
					uDbgListed += __popc ( uCand & uListedRows );
					uDbgHot += __popc ( uCand & ~uListedRows );
					int iOff = __popc ( uCand );
					#pragma unroll
					for ( int d=1; d<32; d<<=1 )
					{
						const int t = __shfl_up_sync ( FULL_MASK, iOff, d );
						if ( iLane>=d )
							iOff += t;
					}
					const int nNew = __shfl_sync ( FULL_MASK, iOff, 31 );
					iOff += nQueue - __popc ( uCand );
					for ( uint32_t m=uCand; m; m&=m-1 )
					{
						const int k = __ffs ( m )-1;
						pQueue[iOff++] = ( uLo + iLane*32 + k ) | ( ( ( uListedRows>>k ) & 1u ) ? OB_LISTED : 0u );
					}
					nQueue += nNew;
					__syncwarp();
					while ( nQueue>=32 )
					{
						nQueue -= 32;
						fnExact ( true, pQueue[nQueue+iLane], pPool, tThrCur );
					}
					__syncwarp();
				}
				++uMini;
			}
			// the queued candidate rows are evaluated against this round's threshold and pool buffer
			if ( nQueue )
			{
				fnExact ( iLane<nQueue, pQueue[iLane<nQueue ? iLane : 0], pPool, tThrCur );
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
					uDbgListed += __shfl_xor_sync ( FULL_MASK, uDbgListed, d );
				}
				if ( iLane==0 )
				{
					atomicAdd ( P.m_pDebug+0, (unsigned long long)uDbgMinis );
					atomicAdd ( P.m_pDebug+1, (unsigned long long)uDbgHot );
					atomicAdd ( P.m_pDebug+2, (unsigned long long)uDbgListed );
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
