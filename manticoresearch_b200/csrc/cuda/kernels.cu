// sm_100a kernels of the full-text query hot path.
//
//   K1  decode_block()       warp-per-skiplist-block VByte doclist decoder
//                            (DiskIndexQword_c::ReadNext, src/sphinx.cpp:511-549; sphUnzipInt src/fileio.cpp:31-45)
//   K2/K3/K6/K7/K8  eval_kernel()  one persistent CTA per work item = (query, rowid range):
//                            decode -> dense-tile boolean evaluation (ExtMultiAnd/ExtAnd/ExtOr/ExtAndNot/ExtMaybe,
//                            src/searchnode.cpp:2570-3711) -> BM25 weight (src/sphinxsearch.cpp:1070, 1096-1141)
//                            -> filters (src/sphinx.cpp:11903) -> threshold-pruned candidate pool + CTA radix select
//                            (CSphMatchQueue semantics, src/sphinxsort.cpp:722-761)
//   K8b merge_kernel()       per query: select + bitonic sort of the work items' candidates, best first
//   K9  shard_merge_kernel() disjoint-rowid-range shard merge (MergeAllMatches, src/searchd.cpp:4653-4738)
//
// All of this is HBM-bound integer/byte work; no tensor cores by design.  Float math uses the
// round-to-nearest intrinsics explicitly (and the file is built with -fmad=false) so that TF*IDF sums
// are bit-identical to the reference's scalar fp32 code.
#include "device_types.h"
#include "hit_stage.cuh"

#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>

namespace mgpu
{

#define FULL_MASK 0xffffffffu

//////////////////////////////////////////////////////////////////////////
// small helpers
//////////////////////////////////////////////////////////////////////////

__device__ __forceinline__ uint4 LdNc16 ( const uint8_t * p )
{
	return __ldg ( reinterpret_cast<const uint4 *>( p ) );
}

__device__ __forceinline__ int WarpInclusiveScan ( int v, int iLane )
{
	#pragma unroll
	for ( int d=1; d<32; d<<=1 )
	{
		int n = __shfl_up_sync ( FULL_MASK, v, d );
		if ( iLane>=d )
			v += n;
	}
	return v;
}

__device__ __forceinline__ uint32_t WarpInclusiveScanU32 ( uint32_t v, int iLane )
{
	#pragma unroll
	for ( int d=1; d<32; d<<=1 )
	{
		uint32_t n = __shfl_up_sync ( FULL_MASK, v, d );
		if ( iLane>=d )
			v += n;
	}
	return v;
}

__device__ __forceinline__ uint64_t WarpInclusiveScanU64 ( uint64_t v, int iLane )
{
	#pragma unroll
	for ( int d=1; d<32; d<<=1 )
	{
		uint64_t n = __shfl_up_sync ( FULL_MASK, v, d );
		if ( iLane>=d )
			v += n;
	}
	return v;
}

__device__ __forceinline__ bool KeyLess ( const Key128_t & a, const Key128_t & b )
{
	return a.m_uHi<b.m_uHi || ( a.m_uHi==b.m_uHi && a.m_uLo<b.m_uLo );
}

/// first index in [uFrom,n) with a[idx]>=x (n if none). Warp-cooperative: a few forward 32-wide probes
/// (consecutive tiles move forward by a handful of blocks), then a uniform binary search.
__device__ uint32_t WarpLowerBound ( const uint32_t * __restrict__ a, uint32_t uFrom, uint32_t n, uint32_t x, int iLane )
{
	uint32_t uPos = uFrom;
	#pragma unroll 1
	for ( int it=0; it<3; ++it )
	{
		uint32_t i = uPos+iLane;
		bool bGe = ( i>=n ) || ( __ldg ( a+i )>=x );
		unsigned m = __ballot_sync ( FULL_MASK, bGe );
		if ( m )
		{
			uint32_t r = uPos + __ffs ( m ) - 1;
			return r<n ? r : n;
		}
		uPos += 32;
	}
	uint32_t lo = uPos, hi = n;
	while ( lo<hi )
	{
		uint32_t mid = lo + ( ( hi-lo )>>1 );
		if ( __ldg ( a+mid )>=x ) hi = mid; else lo = mid+1;
	}
	return lo;
}

//////////////////////////////////////////////////////////////////////////
// K1: warp-cooperative doclist block decoder
//////////////////////////////////////////////////////////////////////////

struct DecodedDoc_t
{
	uint32_t	m_uRowid;
	uint32_t	m_uHits;
	uint32_t	m_uFields;
	uint64_t	m_uHitlistPos;
	bool		m_bValid;
};

/// Decodes block uBlk (<=32 docs) of a leaf's doclist. One doc per lane.
/// pStage: STAGE_BYTES of warp-private shared memory (16 B aligned); pRecStart: 34 x u16.
/// A doclist record is exactly 4 varints (src/sphinx.cpp:8425-8497), so record d starts right after the
/// (4d)-th varint terminator byte of the block: terminators are found with a ballot-free word scan
/// (popc + warp prefix sum), record starts are scattered to shared memory, then each lane decodes its
/// own 4 varints and two warp scans rebuild rowids and hitlist offsets from their deltas.
template<bool NEED_HITPOS>
__device__ __forceinline__ void DecodeBlock ( const DevIndex_t & tIdx, const DevLeaf_t & tLeaf, uint32_t uBlk,
	uint8_t * pStage, uint16_t * pRecStart, int iLane, DecodedDoc_t & tOut )
{
	const uint32_t uGlobalBlk = tLeaf.m_uFirstBlk + uBlk;
	const bool bLast = ( uBlk+1>=tLeaf.m_nBlocks );
	const uint64_t uOff0 = __ldg ( tIdx.m_pBlkOff+uGlobalBlk );
	const uint64_t uOff1 = bLast ? tLeaf.m_uDoclistEnd : __ldg ( tIdx.m_pBlkOff+uGlobalBlk+1 );
	const int nDocs = bLast ? (int)( tLeaf.m_nDocs - 32u*uBlk ) : 32;
	const uint32_t uBaseRowid = __ldg ( tIdx.m_pBlkRowid+uGlobalBlk );

	// stage the block's bytes: 128-bit coalesced loads from a 16 B aligned start
	const uint64_t uAligned = uOff0 & ~15ull;
	const int iHead = (int)( uOff0-uAligned );
	int iTotal = iHead + (int)( uOff1-uOff0 );
	if ( iTotal>STAGE_BYTES-32 )
		iTotal = STAGE_BYTES-32;	// cannot happen for well-formed 32-doc blocks (<= 15+32*25 B); guards corrupt input
	for ( int c=iLane; c*16<iTotal; c+=32 )
		*reinterpret_cast<uint4 *>( pStage+16*c ) = LdNc16 ( tIdx.m_pSpd+uAligned+16*c );
	if ( iLane==0 )
	{
		pRecStart[0] = (uint16_t)iHead;
		// 16 zero bytes behind the staged ones: the per-lane varint loops below need no end test, a corrupt block (continuation
		// bits up to its end) stops here at the latest
		*reinterpret_cast<uint4 *>( pStage + ( ( iTotal+15 ) & ~15 ) ) = make_uint4 ( 0u, 0u, 0u, 0u );
	}
	__syncwarp();

	// find record starts
	int iCarry = 0;
	for ( int iBase=0; iBase<iTotal; iBase+=128 )
	{
		const int iPos = iBase + iLane*4;
		uint32_t w = *reinterpret_cast<const uint32_t *>( pStage+iPos );
		uint32_t t = ( ~w>>7 ) & 0x01010101u;
		// keep only bytes inside [iHead, iTotal)
		int iLo = iHead-iPos, iHi = iTotal-iPos;
		uint32_t m = 0x01010101u;
		if ( iLo>0 )
			m = iLo>=4 ? 0u : ( m & ~( ( 1u<<( 8*iLo ) )-1u ) );
		if ( iHi<4 )
			m = iHi<=0 ? 0u : ( m & ( ( 1u<<( 8*iHi ) )-1u ) );
		t &= m;
		const int c = __popc ( t );
		const int iIncl = WarpInclusiveScan ( c, iLane );
		int iRun = iIncl-c+iCarry;
		#pragma unroll
		for ( int j=0; j<4; ++j )
			if ( ( t>>( 8*j ) ) & 1u )
			{
				if ( ( iRun & 3 )==3 && ( iRun>>2 )<32 )
					pRecStart[( iRun>>2 )+1] = (uint16_t)( iPos+j+1 );
				++iRun;
			}
		iCarry += __shfl_sync ( FULL_MASK, iIncl, 31 );
	}
	__syncwarp();

	// each lane decodes its record
	const bool bValid = iLane<nDocs;
	uint32_t v0 = 0, v1 = 0, v2 = 0;
	uint64_t v3 = 0;
	if ( bValid )
	{
		const uint8_t * p = pStage + pRecStart[iLane];
		uint32_t b;
		do { b = *p++; v0 = ( v0<<7 ) + ( b & 0x7f ); } while ( b & 0x80 );
		if ( tIdx.m_bInlineHits )
		{
			do { b = *p++; v1 = ( v1<<7 ) + ( b & 0x7f ); } while ( b & 0x80 );
			do { b = *p++; v2 = ( v2<<7 ) + ( b & 0x7f ); } while ( b & 0x80 );
			do { b = *p++; v3 = ( v3<<7 ) + ( b & 0x7f ); } while ( b & 0x80 );
		} else
		{
			// plain format: rowid delta, hitlist offset delta, field mask, hits (src/sphinx.cpp:539-545)
			do { b = *p++; v3 = ( v3<<7 ) + ( b & 0x7f ); } while ( b & 0x80 );
			do { b = *p++; v2 = ( v2<<7 ) + ( b & 0x7f ); } while ( b & 0x80 );
			do { b = *p++; v1 = ( v1<<7 ) + ( b & 0x7f ); } while ( b & 0x80 );
		}
	}
	__syncwarp();	// staging buffer may be reused by the caller's next block

	// rowid = (base-1) + prefix sum of deltas, mod 2^32 (reader starts at INVALID_ROWID for block 0)
	tOut.m_uRowid = uBaseRowid - 1u + WarpInclusiveScanU32 ( v0, iLane );
	tOut.m_uHits = v1;
	tOut.m_bValid = bValid;

	const bool bInlined = tIdx.m_bInlineHits && v1==1;
	if ( bInlined )
	{
		// the only hit lives in the doclist: v2 = pos (23 bits), v3 = field<<1 | end (src/sphinx.cpp:523-530)
		uint32_t uField = ( (uint32_t)v3>>1 ) & 255u;
		tOut.m_uFields = uField<32 ? ( 1u<<uField ) : 0u;
	} else
		tOut.m_uFields = v2;

	if ( NEED_HITPOS )
	{
		const uint64_t uBaseHitpos = __ldg ( tIdx.m_pBlkHitpos+uGlobalBlk );
		uint64_t uDelta = ( bValid && !bInlined ) ? v3 : 0ull;
		uint64_t uPos = uBaseHitpos + WarpInclusiveScanU64 ( uDelta, iLane );
		tOut.m_uHitlistPos = bInlined ? ( (uint64_t)v2 | ( v3<<23 ) | ( 1ull<<63 ) ) : uPos;
	} else
		tOut.m_uHitlistPos = 0;
}


/// standalone decode of one doclist (parity tests of K1, and the decode-only roofline probe)
__global__ void __launch_bounds__ ( EVAL_THREADS ) decode_doclist_kernel ( DevIndex_t tIdx, DevLeaf_t tLeaf,
	uint32_t * pRowid, uint32_t * pHits, uint32_t * pFields, uint64_t * pHitlistPos, unsigned long long * pChecksum )
{
	__shared__ __align__(16) uint8_t dStage[EVAL_WARPS][STAGE_BYTES];
	__shared__ uint16_t dRecStart[EVAL_WARPS][34];
	const int iWarp = threadIdx.x>>5, iLane = threadIdx.x & 31;
	unsigned long long uSum = 0;
	for ( uint32_t b = blockIdx.x*EVAL_WARPS+iWarp; b<tLeaf.m_nBlocks; b += gridDim.x*EVAL_WARPS )
	{
		DecodedDoc_t d;
		DecodeBlock<true> ( tIdx, tLeaf, b, dStage[iWarp], dRecStart[iWarp], iLane, d );
		if ( d.m_bValid )
		{
			if ( pRowid )
			{
				size_t i = (size_t)b*32+iLane;
				pRowid[i] = d.m_uRowid;
				pHits[i] = d.m_uHits;
				pFields[i] = d.m_uFields;
				pHitlistPos[i] = d.m_uHitlistPos;
			}
			uSum += d.m_uRowid + d.m_uHits + d.m_uFields + d.m_uHitlistPos;
		}
	}
	if ( pChecksum )
	{
		#pragma unroll
		for ( int d=16; d; d>>=1 )
			uSum += __shfl_xor_sync ( FULL_MASK, uSum, d );
		if ( iLane==0 && uSum )
			atomicAdd ( pChecksum, uSum );
	}
}

//////////////////////////////////////////////////////////////////////////
// CTA-wide exact top-K selection on 128-bit keys (MSB-first radix select)
//////////////////////////////////////////////////////////////////////////

struct SelectSmem_t
{
	uint32_t	m_dHist[256];
	uint64_t	m_uPrefixHi, m_uPrefixLo;
	int			m_iK;
	int			m_iOut;
};

__device__ __forceinline__ bool MatchesPrefix ( const Key128_t & k, uint64_t uPHi, uint64_t uPLo, int iByte )
{
	// true if all bytes above iByte (15 = most significant of hi) equal the prefix
	if ( iByte>=8 )
	{
		int s = 8*( iByte-8 )+8;
		return s>=64 ? true : ( ( k.m_uHi>>s )==( uPHi>>s ) );
	}
	if ( k.m_uHi!=uPHi )
		return false;
	int s = 8*iByte+8;
	return s>=64 ? true : ( ( k.m_uLo>>s )==( uPLo>>s ) );
}

__device__ __forceinline__ uint32_t KeyByte ( const Key128_t & k, int iByte )
{
	return iByte>=8 ? (uint32_t)( ( k.m_uHi>>( 8*( iByte-8 ) ) ) & 255u ) : (uint32_t)( ( k.m_uLo>>( 8*iByte ) ) & 255u );
}

/// Keeps the iK largest of pIn[0..n) (n>iK, keys distinct) in pOut[0..iK); returns the iK-th largest key.
/// Must be called by all threads of the CTA.
__device__ Key128_t CtaSelectTopK ( const Key128_t * pIn, int n, int iK, Key128_t * pOut, SelectSmem_t & s )
{
	const int tid = threadIdx.x, nThreads = blockDim.x;
	if ( tid==0 )
	{
		s.m_uPrefixHi = 0; s.m_uPrefixLo = 0; s.m_iK = iK; s.m_iOut = 0;
	}
	__syncthreads();
	for ( int iByte=15; iByte>=0; --iByte )
	{
		for ( int i=tid; i<256; i+=nThreads )
			s.m_dHist[i] = 0;
		__syncthreads();
		const uint64_t uPHi = s.m_uPrefixHi, uPLo = s.m_uPrefixLo;
		for ( int i=tid; i<n; i+=nThreads )
		{
			Key128_t k = pIn[i];
			if ( MatchesPrefix ( k, uPHi, uPLo, iByte ) )
				atomicAdd ( &s.m_dHist[KeyByte ( k, iByte )], 1u );
		}
		__syncthreads();
		if ( tid<32 )
		{
			// warp 0 finds the digit holding the k-th largest: scan bins from 255 down, 8 bins per lane
			int k = s.m_iK;
			uint32_t uMine = 0;
			#pragma unroll
			for ( int j=0; j<8; ++j )
				uMine += s.m_dHist[255-( tid*8+j )];
			uint32_t uIncl = WarpInclusiveScanU32 ( uMine, tid );
			unsigned m = __ballot_sync ( FULL_MASK, (int)uIncl>=k );
			int iSel = __ffs ( m )-1;		// always found: the prefix bucket holds >= k keys
			if ( tid==iSel )
			{
				int iCum = (int)( uIncl-uMine );
				int d = 255-tid*8;
				for ( int j=0; j<8; ++j, --d )
				{
					int c = (int)s.m_dHist[d];
					if ( iCum+c>=k )
						break;
					iCum += c;
				}
				s.m_iK = k-iCum;
				if ( iByte>=8 )
					s.m_uPrefixHi |= (uint64_t)d<<( 8*( iByte-8 ) );
				else
					s.m_uPrefixLo |= (uint64_t)d<<( 8*iByte );
			}
		}
		__syncthreads();
	}
	Key128_t tThr { s.m_uPrefixHi, s.m_uPrefixLo };
	for ( int i=tid; i<n; i+=nThreads )
	{
		Key128_t k = pIn[i];
		if ( !KeyLess ( k, tThr ) )
		{
			int o = atomicAdd ( &s.m_iOut, 1 );
			if ( o<iK )
				pOut[o] = k;
		}
	}
	__syncthreads();
	return tThr;
}

//////////////////////////////////////////////////////////////////////////
// the fused evaluation kernel
//////////////////////////////////////////////////////////////////////////

struct EvalShared_t
{
	DevQuery_t		m_tQ;
	SelectSmem_t	m_tSel;
	uint32_t		m_dLeafB0[MAX_LEAVES];
	uint32_t		m_dLeafB1[MAX_LEAVES];
	uint32_t		m_dLeafCursor[MAX_LEAVES];
	Key128_t		m_tThr;
	unsigned long long m_uTotal;
	int				m_iItem;
	int				m_iPoolCnt;
	int				m_iPoolBuf;
	int				m_iListCnt;
	uint32_t		m_uNextRow;
	int				m_dPreOff[MAX_LEAVES+1];	///< first predecode block slot of each leaf in this tile
	uint32_t		m_dRankTab[16];		///< ExtRanker_WeightSum_c: sum of the weights of the fields in a 4-bit mask (indexes with <= 4 fields)
	uint32_t		m_dAliveBits[TILE_W/32];
	// quorum node: the children vector's order per rowid interval (ExtQuorum_c removes a keyword with RemoveFast when it runs out of documents)
	uint32_t		m_dQLast[MAX_PHRASE_WORDS];								///< last rowid of each child (0xFFFFFFFF: no documents)
	uint32_t		m_dQBound[MAX_PHRASE_WORDS+1];							///< interval k holds the rows <= m_dQBound[k]
	uint8_t			m_dQOrder[MAX_PHRASE_WORDS+1][MAX_PHRASE_WORDS];
	uint8_t			m_dQLen[MAX_PHRASE_WORDS+1];
	int				m_nQIntervals;
	float			m_dTf[256];			///< float(hits)/float(hits+1.2f), src/searchnode.cpp:1946
	uint16_t		m_dRecStart[EVAL_WARPS][34];
	__align__(16) uint8_t m_dStage[EVAL_WARPS][STAGE_BYTES];
};

/// dense doc vectors live in dynamic shared memory: [nStack] x { float tfidf[W]; u32 fields[W]; u8 cnt[W] }
/// hit stage adds { u32 emit[W] } per level (which keywords / n-way nodes sit on the doc: the CollectHits() recursion,
/// src/searchnode.cpp:2627-2705, 3516-3545) and one u16 list of compacted candidate slots
struct Vectors_t
{
	float *		m_pTfidf;
	uint32_t *	m_pFields;
	uint8_t *	m_pCnt;
	uint32_t *	m_pEmit;
	uint16_t *	m_pList;
	__device__ __forceinline__ float &		Tfidf ( int v, int s )	{ return m_pTfidf[v*TILE_W+s]; }
	__device__ __forceinline__ uint32_t &	Fields ( int v, int s )	{ return m_pFields[v*TILE_W+s]; }
	__device__ __forceinline__ uint8_t &	Cnt ( int v, int s )	{ return m_pCnt[v*TILE_W+s]; }
	__device__ __forceinline__ uint32_t &	Emit ( int v, int s )	{ return m_pEmit[v*TILE_W+s]; }
};

/// CTA-wide compaction of the slots of v[iVec] whose cnt equals uAlive into V.m_pList; returns the count
__device__ int CompactAlive ( Vectors_t & V, int iVec, uint8_t uAlive, int nSlots, int * pCount )
{
	const int tid = threadIdx.x, iLane = tid & 31;
	if ( tid==0 )
		*pCount = 0;
	__syncthreads();
	for ( int sBase=0; sBase<nSlots; sBase+=EVAL_THREADS )
	{
		const int s = sBase+tid;
		const bool b = s<nSlots && V.Cnt ( iVec, s )==uAlive;
		const unsigned m = __ballot_sync ( FULL_MASK, b );
		if ( m )
		{
			int iBase = 0;
			if ( iLane==0 )
				iBase = atomicAdd ( pCount, __popc ( m ) );
			iBase = __shfl_sync ( FULL_MASK, iBase, 0 );
			if ( b )
				V.m_pList[iBase + __popc ( m & ( ( 1u<<iLane )-1u ) )] = (uint16_t)s;
		}
	}
	__syncthreads();
	return *pCount;
}

__device__ __forceinline__ int64_t LoadAttr ( const DevIndex_t & tIdx, uint32_t uRowid, int iDwordOff, int iBitCount )
{
	const uint32_t * p = tIdx.m_pSpa + (size_t)uRowid*tIdx.m_iStride + iDwordOff;
	uint32_t lo = __ldg ( p );
	if ( iBitCount==64 )
		return (int64_t)( (uint64_t)lo | ( (uint64_t)__ldg ( p+1 )<<32 ) );
	return (int64_t)lo;
}

__device__ __forceinline__ bool PassFilters ( const DevIndex_t & tIdx, const DevQuery_t & q, uint32_t uRowid )
{
	for ( int f=0; f<q.m_nFilters; ++f )
	{
		const DevFilter_t & t = q.m_dFilters[f];
		int64_t v = LoadAttr ( tIdx, uRowid, t.m_iDwordOff, t.m_iBitCount );
		bool bOk;
		if ( t.m_eKind==0 )
			bOk = ( v>=t.m_iMin && v<=t.m_iMax );
		else
		{
			bOk = false;
			for ( int k=0; k<t.m_nValues; ++k )
				bOk |= ( t.m_dValues[k]==v );
		}
		if ( t.m_bExclude )
			bOk = !bOk;
		if ( !bOk )
			return false;
	}
	return true;
}

/// packs the sort keys of one match: bigger key = better match (CSphMatchComparatorState semantics)
__device__ __forceinline__ Key128_t MakeKey ( const DevIndex_t & tIdx, const DevQuery_t & q, uint32_t uRowid, int iWeight )
{
	Key128_t k;
	const uint32_t uW = (uint32_t)iWeight ^ 0x80000000u;
	if ( q.m_nSortKeys==0 )
		k.m_uHi = (uint64_t)uW<<32;	// MatchRelevanceLt_fn
	else
	{
		uint64_t uHi = 0;
		for ( int i=0; i<q.m_nSortKeys; ++i )
		{
			const DevSortKey_t & t = q.m_dSortKeys[i];
			uint64_t v;
			int iBits = 32;
			if ( t.m_eKind==1 )			v = uW;
			else if ( t.m_eKind==0 )	v = uRowid + tIdx.m_uRowidBase;
			else if ( t.m_eKind==3 )
			{
				// SPH_KEYPART_FLOAT: IEEE bits -> an unsigned key of the same order (-0 joins +0)
				uint32_t u = (uint32_t)LoadAttr ( tIdx, uRowid, t.m_iDwordOff, 32 );
				if ( u==0x80000000u )
					u = 0;
				v = ( u & 0x80000000u ) ? (uint32_t)~u : ( u | 0x80000000u );
			} else
			{
				int64_t a = LoadAttr ( tIdx, uRowid, t.m_iDwordOff, t.m_iBitCount );
				if ( t.m_iBitCount==64 ) { v = (uint64_t)a ^ 0x8000000000000000ull; iBits = 64; }
				else v = (uint64_t)a;
			}
			if ( !t.m_bDesc )
				v = iBits==64 ? ~v : ( ~v & 0xffffffffull );
			uHi |= v<<t.m_iShift;
		}
		k.m_uHi = uHi;
	}
	k.m_uLo = ( (uint64_t)( ~( uRowid+tIdx.m_uRowidBase ) )<<32 ) | (uint32_t)iWeight;
	return k;
}

/// one posting of a keyword meets tile slot s of v[d]: the per-document step of ExtTerm_T (SET), ExtAnd_c / ExtMultiAnd_T (AND),
/// ExtOr_c (OR), ExtAndNot_c (ANDNOT), ExtMaybe_c (MAYBE). Returns whether the keyword "sits on" the doc for CollectHits().
template<bool HITS, typename VEC>
__device__ __forceinline__ bool ApplyTermOp ( VEC & V, const DevOp_t & tOp, int d, int s, float fTf, uint32_t uFields, uint32_t uEmitBit )
{
	const uint8_t uAlive = tOp.m_uAliveDst;
	switch ( tOp.m_eCode )
	{
	case OP_TERM_SET:
		V.Tfidf ( d, s ) = fTf; V.Fields ( d, s ) = uFields; V.Cnt ( d, s ) = 1;
		if ( HITS ) V.Emit ( d, s ) = uEmitBit;
		return true;
	case OP_TERM_AND:
		if ( V.Cnt ( d, s )==uAlive )
		{
			V.Tfidf ( d, s ) = __fadd_rn ( V.Tfidf ( d, s ), fTf );
			V.Fields ( d, s ) |= uFields;
			V.Cnt ( d, s ) = tOp.m_uAliveOut;
			if ( HITS ) V.Emit ( d, s ) |= uEmitBit;
			return true;
		}
		return false;
	case OP_TERM_OR:
		if ( V.Cnt ( d, s )==uAlive )
		{
			V.Tfidf ( d, s ) = __fadd_rn ( V.Tfidf ( d, s ), fTf );
			V.Fields ( d, s ) |= uFields;
			if ( HITS ) V.Emit ( d, s ) |= uEmitBit;
		} else
		{
			V.Tfidf ( d, s ) = fTf; V.Fields ( d, s ) = uFields; V.Cnt ( d, s ) = uAlive;
			if ( HITS ) V.Emit ( d, s ) = uEmitBit;
		}
		return true;
	case OP_TERM_ANDNOT:
		if ( V.Cnt ( d, s )==uAlive )
			V.Cnt ( d, s ) = 0;
		return false;
	default: // OP_TERM_MAYBE
		if ( V.Cnt ( d, s )==uAlive )
		{
			V.Tfidf ( d, s ) = __fadd_rn ( V.Tfidf ( d, s ), fTf );
			V.Fields ( d, s ) |= uFields;
			if ( HITS ) V.Emit ( d, s ) |= uEmitBit;
			return true;
		}
		return false;
	}
}

/// warp-uniform: any alive bit in slots [uLo, uHi) of the 2048-bit map
__device__ __forceinline__ bool AnyAliveInRange ( const uint32_t * pBits, uint32_t uLo, uint32_t uHi, int iLane )
{
	if ( uLo>=uHi )
		return false;
	bool bAny = false;
	#pragma unroll
	for ( int k=0; k<TILE_W/32/32; ++k )
	{
		const uint32_t w = (uint32_t)( k*32+iLane );
		const uint32_t uBase = w*32;
		uint32_t m = pBits[w];
		if ( uBase+32<=uLo || uBase>=uHi )
			m = 0;
		else
		{
			if ( uLo>uBase )
				m &= ~0u<<( uLo-uBase );
			if ( uHi<uBase+32 )
				m &= ( 1u<<( uHi-uBase ) )-1u;
		}
		bAny |= ( m!=0 );
	}
	return __any_sync ( FULL_MASK, bAny );
}

/// documents with >= 255 hits of a hot keyword: their hit counts live in a list of { hot slot, rowid, hits, next } entries chained
/// per bucket of a 65536-way hash of ( slot, rowid ); the bucket heads follow the entry counter in m_pEscapeCount. A lookup walks one
/// chain (n / 65536 entries on average), not the list
static const int HOT_ESCAPE_BUCKETS = 65536;
__device__ __forceinline__ uint32_t HotEscapeBucket ( uint32_t uHot, uint32_t uRowid )
{
	return ( ( uRowid*2654435761u ) ^ ( uHot*40503u ) )>>16;
}

__device__ uint32_t HotEscapeHits ( const DevHotStore_t & tHot, int iHot, uint32_t uRowid )
{
	const int nCap = __ldg ( tHot.m_pEscapeCount );
	int i = __ldg ( tHot.m_pEscapeCount+1+HotEscapeBucket ( (uint32_t)iHot, uRowid ) );
	for ( int nSteps=0; i>=0 && nSteps<=nCap; ++nSteps )
	{
		const uint4 t = __ldg ( reinterpret_cast<const uint4 *>( tHot.m_pEscape )+i );
		if ( t.x==(uint32_t)iHot && t.y==uRowid )
			return t.z;
		i = (int)t.w;
	}
	return 255u;
}

/// K0: decodes every hot keyword's doclist ONCE per batch into the dense store (warp per 32-doc block). The blocks of all hot
/// keywords form one flat list (P.m_pBlkStart = prefix sums of the keywords' block counts): a warp takes HOT_CHUNK consecutive
/// blocks at a time, so short doclists keep every warp busy (a 0.5 % keyword on a 1.25M-row shard has 195 blocks).
/// Next to the u16 row {hits, fields, tf class} the kernel sets the row's bit in the keyword's per-field presence bitmaps
/// (orbits_kernel); lanes of a block whose rows share a bitmap word combine their bits before the one atomic per word.
static const int HOT_CHUNK = 8;

__global__ void __launch_bounds__ ( EVAL_THREADS ) hot_decode_kernel ( HotDecodeParams_t P )
{
	__shared__ __align__(16) uint8_t dStage[EVAL_WARPS][STAGE_BYTES];
	__shared__ uint16_t dRecStart[EVAL_WARPS][34];
	const int iWarp = threadIdx.x>>5, iLane = threadIdx.x & 31;
	const uint32_t uWarpGlobal = blockIdx.x*EVAL_WARPS+iWarp, nWarps = gridDim.x*EVAL_WARPS;
	const uint32_t nTotalBlocks = __ldg ( P.m_pBlkStart+P.m_nHot );
	const size_t iBitStride = (size_t)( P.m_iStride>>5 );
	// tf class = ceil ( 15*h/(h+1.2) ) in exact integers: class/15 >= tf, the share of the weight bound of stream_kernel<512,2>
	__shared__ uint8_t dClass[256];
	{
		const uint32_t uH = threadIdx.x & 255u;
		dClass[uH] = P.m_bTfClass ? (uint8_t)( ( 150u*uH + 10u*uH+11u )/( 10u*uH+12u ) ) : 0;
	}
	__syncthreads();
	for ( uint32_t g0=uWarpGlobal*HOT_CHUNK; g0<nTotalBlocks; g0+=nWarps*HOT_CHUNK )
	{
		// the keyword holding flat block g0: last h with start[h] <= g0
		int h = 0;
		{
			int lo = 0, hi = P.m_nHot;
			while ( hi-lo>1 )
			{
				const int mid = ( lo+hi )>>1;
				if ( __ldg ( P.m_pBlkStart+mid )<=g0 ) lo = mid; else hi = mid;
			}
			h = lo;
		}
		const uint32_t g1 = min ( g0+(uint32_t)HOT_CHUNK, nTotalBlocks );
		for ( uint32_t g=g0; g<g1; ++g )
		{
			while ( g>=__ldg ( P.m_pBlkStart+h+1 ) )
				++h;
			const DevLeaf_t tLeaf = P.m_pTerms[h];
			uint16_t * pD = P.m_pData + (size_t)h*P.m_iStride;
			DecodedDoc_t d;
			const uint32_t uBlk = g-__ldg ( P.m_pBlkStart+h );
			DecodeBlock<false> ( P.m_tIndex, tLeaf, uBlk, dStage[iWarp], dRecStart[iWarp], iLane, d );
			const bool bValid = d.m_bValid && d.m_uRowid<P.m_tIndex.m_uRows;	// (a corrupt doclist must not write outside the store)
			{
				// The store is not cleared beforehand: every block zero-fills its own stretch of the keyword's row, from the row after the
				// previous block's last posting (= the block's skiplist base) through its own last posting, with coalesced 16 B stores; the
				// keyword's last block also clears the rest of the row. The postings land on top, still in L2 (25.8 GB of cudaMemset per
				// 10k-query batch gone, and the scattered 2 B writes no longer pull their sectors back from DRAM).
				const unsigned uValidMask = __ballot_sync ( FULL_MASK, bValid );
				const uint32_t uStride = (uint32_t)P.m_iStride;
				uint32_t uFillLo = uBlk ? min ( __ldg ( P.m_tIndex.m_pBlkRowid+tLeaf.m_uFirstBlk+uBlk ), uStride ) : 0u;
				uint32_t uFillHi = uFillLo;
				if ( uValidMask )
					uFillHi = __shfl_sync ( FULL_MASK, d.m_uRowid, 31-__clz ( uValidMask ) )+1u;
				if ( uBlk+1>=tLeaf.m_nBlocks )
					uFillHi = uStride;
				if ( uFillHi>uFillLo )
				{
					const uint32_t uBody0 = min ( ( uFillLo+7u ) & ~7u, uFillHi ), uBody1 = max ( uFillHi & ~7u, uBody0 );
					if ( uFillLo+iLane<uBody0 )
						pD[uFillLo+iLane] = 0;
					uint4 * pVec = reinterpret_cast<uint4 *>( pD );
					for ( uint32_t i=( uBody0>>3 )+iLane; i<( uBody1>>3 ); i+=32 )
						pVec[i] = make_uint4 ( 0u, 0u, 0u, 0u );
					if ( uBody1+iLane<uFillHi )
						pD[uBody1+iLane] = 0;
				}
				__syncwarp();
			}
			if ( bValid )
			{
				const uint32_t uHits = min ( d.m_uHits, 255u );
				pD[d.m_uRowid] = (uint16_t)( uHits | ( ( d.m_uFields & 255u )<<8 ) | ( (uint32_t)dClass[uHits]<<12 ) );
				if ( d.m_uHits>=255u )
				{
					const int i = atomicAdd ( P.m_pEscapeCount, 1 );
					if ( i<P.m_iEscapeCap )
					{
						const int iNext = atomicExch ( P.m_pEscapeCount+1+HotEscapeBucket ( (uint32_t)h, d.m_uRowid ), i );
						reinterpret_cast<uint4 *>( P.m_pEscape )[i] = make_uint4 ( (uint32_t)h, d.m_uRowid, d.m_uHits, (uint32_t)iNext );
					}
				}
			}
			if ( P.m_nBitFields )
			{
				// Rowids ascend inside a block, so the lanes of one 32-row bitmap word are neighbours. Sparse keywords: (nearly) every
				// lane has a word of its own and sets its bits directly. Dense ones: the block spans a few words; per word one
				// full-mask redux per bitmap and one atomic by the run's first lane.
				const uint32_t uWord = bValid ? ( d.m_uRowid>>5 ) : 0xFFFFFFFFu;
				const uint32_t uPrevWord = __shfl_up_sync ( FULL_MASK, uWord, 1 );
				const bool bHead = bValid && ( iLane==0 || uPrevWord!=uWord );
				const unsigned uHeads = __ballot_sync ( FULL_MASK, bHead );
				const uint32_t uBit = 1u<<( d.m_uRowid & 31u );
				const int iLvl = P.m_pLvlSlot ? __ldg ( P.m_pLvlSlot+h ) : -1;
				uint32_t * pB = P.m_pBits + (size_t)h*P.m_nBitFields*iBitStride;
				uint32_t * pL = iLvl>=0 ? P.m_pLvlBits + (size_t)iLvl*2*iBitStride : nullptr;
				if ( __popc ( uHeads )>8 )
				{
					// (a few lanes may share a word: their atomics simply hit it twice)
					if ( bValid )
					{
						for ( int f=0; f<P.m_nBitFields; ++f )
							if ( ( d.m_uFields>>f ) & 1u )
								atomicOr ( pB + f*iBitStride + uWord, uBit );
						if ( pL && d.m_uHits>=2u )
						{
							atomicOr ( pL + uWord, uBit );
							if ( d.m_uHits>=4u )
								atomicOr ( pL + iBitStride + uWord, uBit );
						}
					}
				} else
					for ( unsigned m=uHeads; m; m&=m-1 )
					{
						const int iHeadLane = __ffs ( m )-1;
						const uint32_t uRunWord = __shfl_sync ( FULL_MASK, uWord, iHeadLane );
						const uint32_t uMine = ( bValid && uWord==uRunWord ) ? uBit : 0u;
						for ( int f=0; f<P.m_nBitFields; ++f )
						{
							const uint32_t uAll = __reduce_or_sync ( FULL_MASK, ( ( d.m_uFields>>f ) & 1u ) ? uMine : 0u );
							if ( uAll && iLane==iHeadLane )
								atomicOr ( pB + f*iBitStride + uRunWord, uAll );
						}
						if ( pL )
						{
							const uint32_t uAll2 = __reduce_or_sync ( FULL_MASK, d.m_uHits>=2u ? uMine : 0u );
							const uint32_t uAll4 = __reduce_or_sync ( FULL_MASK, d.m_uHits>=4u ? uMine : 0u );
							if ( iLane==iHeadLane )
							{
								if ( uAll2 )
									atomicOr ( pL + uRunWord, uAll2 );
								if ( uAll4 )
									atomicOr ( pL + iBitStride + uRunWord, uAll4 );
							}
						}
					}
			}
		}
	}
}

/// K0b: decodes the non-hot keywords of launch class 5 ONCE per batch into plain posting lists (rowid, hits | fields<<24).
/// Same flat (keyword, block) walk as K0; block b of a keyword lands at entries [32b, 32b+32) of its list: coalesced stores.
__global__ void __launch_bounds__ ( EVAL_THREADS ) sparse_decode_kernel ( SparseDecodeParams_t P )
{
	__shared__ __align__(16) uint8_t dStage[EVAL_WARPS][STAGE_BYTES];
	__shared__ uint16_t dRecStart[EVAL_WARPS][34];
	const int iWarp = threadIdx.x>>5, iLane = threadIdx.x & 31;
	const uint32_t uWarpGlobal = blockIdx.x*EVAL_WARPS+iWarp, nWarps = gridDim.x*EVAL_WARPS;
	const uint32_t nTotalBlocks = __ldg ( P.m_pBlkStart+P.m_nTerms );
	for ( uint32_t g0=uWarpGlobal*HOT_CHUNK; g0<nTotalBlocks; g0+=nWarps*HOT_CHUNK )
	{
		int h = 0;
		{
			int lo = 0, hi = P.m_nTerms;
			while ( hi-lo>1 )
			{
				const int mid = ( lo+hi )>>1;
				if ( __ldg ( P.m_pBlkStart+mid )<=g0 ) lo = mid; else hi = mid;
			}
			h = lo;
		}
		const uint32_t g1 = min ( g0+(uint32_t)HOT_CHUNK, nTotalBlocks );
		for ( uint32_t g=g0; g<g1; ++g )
		{
			while ( g>=__ldg ( P.m_pBlkStart+h+1 ) )
				++h;
			const DevLeaf_t tLeaf = P.m_pTerms[h];
			const uint32_t b = g-__ldg ( P.m_pBlkStart+h );
			DecodedDoc_t d;
			DecodeBlock<false> ( P.m_tIndex, tLeaf, b, dStage[iWarp], dRecStart[iWarp], iLane, d );
			if ( d.m_bValid )
			{
				const size_t i = (size_t)tLeaf.m_uListOff + 32u*b + iLane;
				P.m_pRows[i] = d.m_uRowid;
				P.m_pVals[i] = min ( d.m_uHits, 0xFFFFFFu ) | ( ( d.m_uFields & 255u )<<24 );
			}
		}
	}
}

template<bool HITS>
__global__ void __launch_bounds__ ( EVAL_THREADS ) eval_kernel ( EvalParams_t P, int nStack )
{
	extern __shared__ __align__(16) uint8_t dDyn[];
	__shared__ EvalShared_t S;

	Vectors_t V;
	V.m_pTfidf = reinterpret_cast<float *>( dDyn );
	V.m_pFields = reinterpret_cast<uint32_t *>( dDyn + (size_t)nStack*TILE_W*4 );
	V.m_pEmit = reinterpret_cast<uint32_t *>( dDyn + (size_t)nStack*TILE_W*8 );
	V.m_pList = reinterpret_cast<uint16_t *>( dDyn + (size_t)nStack*TILE_W*( HITS ? 12 : 8 ) );
	V.m_pCnt = dDyn + (size_t)nStack*TILE_W*( HITS ? 12 : 8 ) + ( HITS ? TILE_W*2 : 0 );
	uint64_t * pHitpos = HITS ? P.m_pHitpos + (size_t)blockIdx.x*MAX_LEAVES*TILE_W : nullptr;
	uint64_t * pLeafTf = HITS ? P.m_pLeafTf + (size_t)blockIdx.x*MAX_LEAVES*TILE_W : nullptr;
	PreEntry_t * pPre = P.m_pPre + (size_t)blockIdx.x*PRE_BLOCKS*32;
	uint64_t * pPreHitpos = HITS ? P.m_pPreHitpos + (size_t)blockIdx.x*PRE_BLOCKS*32 : nullptr;

	const int tid = threadIdx.x, iWarp = tid>>5, iLane = tid & 31;
	const DevIndex_t & tIdx = P.m_tIndex;
	Key128_t * pPool0 = P.m_pPool + (size_t)blockIdx.x*2*P.m_iPoolCap;
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

		// stage the query descriptor
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
		if ( tid<MAX_LEAVES )
			S.m_dLeafCursor[tid] = 0;
		__syncthreads();
		const DevQuery_t & q = S.m_tQ;
		const int iK = q.m_iMaxMatches;
		int iMyTotal = 0;
		unsigned long long uMyHitBytes = 0;

		if ( tid<16 )
		{
			uint32_t uSum = 0;
			for ( int i=0; i<4 && i<q.m_nWeights; ++i )
				if ( tid & ( 1<<i ) )
					uSum += (uint32_t)q.m_dWeights[i];
			S.m_dRankTab[tid] = uSum;
		}
		const bool bJump = q.m_iDriverLeaf>=0;	// pure AND query opened by a sparse keyword: visit only the tiles that keyword touches
		if constexpr ( HITS )
		{
			// a quorum node (the planner admits one per query): the reference sums TF*IDF in the order of its children vector, and
			// RemoveFast moves the last child into the slot of every keyword that has run out of documents. The order of each
			// rowid interval follows from the keywords' last rowids.
			int jq = -1;
			for ( int j=0; j<q.m_nNWay; ++j )
				if ( q.m_dNWay[j].m_eKind==NWAY_QUORUM )
					jq = j;
			if ( jq>=0 )
			{
				const DevNWay_t & tN = q.m_dNWay[jq];
				if ( iWarp==0 )
					for ( int w=0; w<tN.m_nWords; ++w )
					{
						const DevLeaf_t & tLeaf = q.m_dLeaves[tN.m_dLeaf[w]];
						uint32_t uLast = 0xFFFFFFFFu;
						if ( tLeaf.m_nBlocks )
						{
							DecodedDoc_t tDoc;
							DecodeBlock<false> ( tIdx, tLeaf, tLeaf.m_nBlocks-1, S.m_dStage[0], S.m_dRecStart[0], iLane, tDoc );
							uLast = __shfl_sync ( FULL_MASK, tDoc.m_uRowid, (int)( tLeaf.m_nDocs-32u*( tLeaf.m_nBlocks-1 ) )-1 );
						}
						if ( iLane==0 )
							S.m_dQLast[w] = uLast;
						__syncwarp();
					}
				__syncthreads();
				if ( tid==0 )
				{
					uint8_t dVec[MAX_PHRASE_WORDS];
					int nLen = tN.m_nWords, nInt = 0;
					for ( int i=0; i<nLen; ++i )
						dVec[i] = (uint8_t)i;
					// warm-up (:4472-4484): children without documents go first
					for ( int i=0; i<nLen; ++i )
						if ( S.m_dQLast[dVec[i]]==0xFFFFFFFFu )
						{
							dVec[i] = dVec[--nLen];
							--i;
						}
					while ( nLen>0 && nInt<MAX_PHRASE_WORDS+1 )
					{
						uint32_t uMin = 0xFFFFFFFFu;
						for ( int i=0; i<nLen; ++i )
							uMin = min ( uMin, S.m_dQLast[dVec[i]] );
						S.m_dQBound[nInt] = uMin;
						S.m_dQLen[nInt] = (uint8_t)nLen;
						for ( int i=0; i<nLen; ++i )
							S.m_dQOrder[nInt][i] = dVec[i];
						++nInt;
						for ( int i=0; i<nLen; ++i )
							if ( S.m_dQLast[dVec[i]]==uMin )
							{
								dVec[i] = dVec[--nLen];
								--i;
							}
					}
					if ( !nInt )
					{
						S.m_dQBound[0] = 0xFFFFFFFFu;
						S.m_dQLen[0] = 0;
						nInt = 1;
					}
					S.m_nQIntervals = nInt;
				}
				__syncthreads();
			}
		}

		uint32_t uTileLo = tItem.m_uRowLo;
		while ( uTileLo<tItem.m_uRowHi )
		{
			const uint32_t uTileHi = min ( uTileLo+(uint32_t)TILE_W, tItem.m_uRowHi );

			// compact the candidate pool if this tile could overflow it
			if ( S.m_iPoolCnt+TILE_W>P.m_iPoolCap )
			{
				Key128_t * pIn = pPool0 + (size_t)S.m_iPoolBuf*P.m_iPoolCap;
				Key128_t * pOut = pPool0 + (size_t)( S.m_iPoolBuf^1 )*P.m_iPoolCap;
				Key128_t tThr = CtaSelectTopK ( pIn, S.m_iPoolCnt, iK, pOut, S.m_tSel );
				if ( tid==0 )
				{
					S.m_tThr = tThr;
					S.m_iPoolCnt = iK;
					S.m_iPoolBuf ^= 1;
				}
				__syncthreads();
			}

			// block range of every sparse leaf inside this tile
			for ( int l=iWarp; l<q.m_nLeaves; l+=EVAL_WARPS )
			{
				const DevLeaf_t & tLeaf = q.m_dLeaves[l];
				uint32_t b0 = 0, b1 = 0;
				if ( tLeaf.m_nBlocks && ( HITS || tLeaf.m_iHot<0 ) )
				{
					const uint32_t * pBase = tIdx.m_pBlkRowid + tLeaf.m_uFirstBlk;
					// b0 = last block whose base <= tile lo; b1 = first block whose base >= tile hi
					uint32_t u = WarpLowerBound ( pBase, S.m_dLeafCursor[l], tLeaf.m_nBlocks, uTileLo+1, iLane );
					b0 = u ? u-1 : 0;
					b1 = WarpLowerBound ( pBase, u, tLeaf.m_nBlocks, uTileHi, iLane );
				}
				if ( iLane==0 )
				{
					S.m_dLeafB0[l] = b0;
					S.m_dLeafB1[l] = b1;
					S.m_dLeafCursor[l] = b1 ? b1-1 : 0;
				}
			}
			if ( tid==0 )
				S.m_uNextRow = 0xFFFFFFFFu;
			__syncthreads();

			// Tile predecode: every block of every scattering (non-chain) sparse keyword that overlaps the tile is decoded in ONE
			// cooperative phase -- all warps busy -- into this CTA's scratch list; the keyword's op then scatters its entries with
			// the whole CTA. (AND-chain keywords stay lazy: they only decode blocks that still hold candidates.)
			if ( tid==0 )
			{
				int iOff = 0;
				for ( int l=0; l<q.m_nLeaves; ++l )
				{
					S.m_dPreOff[l] = iOff;
					if ( ( ( q.m_uPreMask>>l ) & 1u ) && ( HITS || q.m_dLeaves[l].m_iHot<0 ) )
						iOff += (int)( S.m_dLeafB1[l]-S.m_dLeafB0[l] );
				}
				S.m_dPreOff[q.m_nLeaves] = iOff;
			}
			__syncthreads();
			{
				const int nPre = S.m_dPreOff[q.m_nLeaves];
				for ( int p=iWarp; p<nPre; p+=EVAL_WARPS )
				{
					int l = 0;
					while ( S.m_dPreOff[l+1]<=p )
						++l;
					const DevLeaf_t & tLeaf = q.m_dLeaves[l];
					const uint32_t b = S.m_dLeafB0[l] + (uint32_t)( p-S.m_dPreOff[l] );
					DecodedDoc_t tDoc;
					DecodeBlock<HITS> ( tIdx, tLeaf, b, S.m_dStage[iWarp], S.m_dRecStart[iWarp], iLane, tDoc );
					const uint32_t uFields = tDoc.m_uFields & tLeaf.m_uQueriedFields;
					bool bOk = tDoc.m_bValid && uFields;
					if ( HITS && bOk && tLeaf.m_iTermPos && tDoc.m_uRowid>=uTileLo && tDoc.m_uRowid<uTileHi )
						bOk = HasAcceptableHit ( tIdx.m_pSpp, tDoc.m_uHitlistPos, tLeaf.m_uQueriedFields, tLeaf.m_iTermPos );
					if ( bOk && tDoc.m_uRowid>=uTileHi && q.m_iDriverLeaf==l )
						atomicMin ( &S.m_uNextRow, tDoc.m_uRowid );	// exact next candidate of a pure AND query: tiles in between are skipped
					PreEntry_t tEntry;
					tEntry.m_uRowid = ( bOk && tDoc.m_uRowid>=uTileLo && tDoc.m_uRowid<uTileHi ) ? tDoc.m_uRowid : 0xFFFFFFFFu;
					// ExtTerm_T::GetDocsChunk, src/searchnode.cpp:1946
					const float fHits = __uint2float_rn ( tDoc.m_uHits );
					tEntry.m_fTf = __fmul_rn ( __fdiv_rn ( fHits, __fadd_rn ( fHits, 1.2f ) ), tLeaf.m_fIDF );
					tEntry.m_uFields = uFields;
					tEntry.m_uPad = 0;
					pPre[(size_t)p*32+iLane] = tEntry;
					if ( HITS )
						pPreHitpos[(size_t)p*32+iLane] = tDoc.m_uHitlistPos;
				}
			}
			__syncthreads();

			// Run the tile program. Dense keyword ops and vector ops are slot-local: thread t only ever touches slots
			// t, t+256, ... so they need no barrier between them; barriers surround the scattering (sparse) ops only.
			bool bDirty = false;	// slot-local writes not yet published by a barrier
			for ( int iOp=0; iOp<q.m_nOps; ++iOp )
			{
				const DevOp_t tOp = q.m_dOps[iOp];
				const int d = tOp.m_uDst;
				if ( tOp.m_eCode<=OP_TERM_MAYBE )
				{
					const DevLeaf_t & tLeaf = q.m_dLeaves[tOp.m_uLeaf];
					const bool bDense = !HITS && tLeaf.m_iHot>=0;
					const uint32_t b0 = S.m_dLeafB0[tOp.m_uLeaf], b1 = S.m_dLeafB1[tOp.m_uLeaf];

					// AND chains: no candidate left in the tile (or an opening keyword without postings here) -> resume after the chain
					const bool bChain = ( tOp.m_uSrc!=0 );
					if ( bChain && tOp.m_eCode==OP_TERM_SET && !bDense && b0>=b1 )
					{
						#pragma unroll
						for ( int k=0; k<TILE_W/EVAL_THREADS; ++k )
							V.Cnt ( d, k*EVAL_THREADS+tid ) = 0;	// slot-local: no barrier needed
						bDirty = true;
						iOp = (int)tOp.m_uSrc-2;
						continue;
					}
					if ( bChain && tOp.m_eCode==OP_TERM_AND )
					{
						bool bMine = false;
						#pragma unroll
						for ( int k=0; k<TILE_W/EVAL_THREADS; ++k )
						{
							const int s = k*EVAL_THREADS+tid;
							const bool bA = V.Cnt ( d, s )==tOp.m_uAliveDst;
							if ( !bDense )
							{
								const unsigned m = __ballot_sync ( FULL_MASK, bA );
								if ( iLane==0 )
									S.m_dAliveBits[s>>5] = m;
							}
							bMine |= bA;
						}
						const int iAny = __syncthreads_or ( bMine );
						bDirty = false;
						if ( !iAny )
						{
							iOp = (int)tOp.m_uSrc-2;
							continue;
						}
					}

					if ( bDense )
					{
						// the keyword comes from the batch's dense hot-term store: u16 per row = hits | fields<<8.
						// All 8 loads of a thread are independent and issued up front (one round trip per op).
						const uint16_t * pD = P.m_tHot.m_pData + (size_t)tLeaf.m_iHot*P.m_tHot.m_iStride + uTileLo;
						uint32_t dRaw[TILE_W/EVAL_THREADS];
						#pragma unroll
						for ( int k=0; k<TILE_W/EVAL_THREADS; ++k )
							dRaw[k] = __ldg ( pD + k*EVAL_THREADS+tid );
						#pragma unroll
						for ( int k=0; k<TILE_W/EVAL_THREADS; ++k )
						{
							const int s = k*EVAL_THREADS+tid;
							uint32_t uHits = dRaw[k] & 255u;
							const uint32_t uFields = uHits ? ( ( dRaw[k]>>8 ) & tLeaf.m_uQueriedFields ) : 0u;
							if ( !uFields )
							{
								if ( tOp.m_eCode==OP_TERM_SET )
									V.Cnt ( d, s ) = 0;
								continue;
							}
							float fBase;
							if ( uHits<255 )
								fBase = S.m_dTf[uHits];
							else
							{
								uHits = HotEscapeHits ( P.m_tHot, tLeaf.m_iHot, uTileLo+s );
								const float fHits = __uint2float_rn ( uHits );
								fBase = __fdiv_rn ( fHits, __fadd_rn ( fHits, 1.2f ) );
							}
							ApplyTermOp<false> ( V, tOp, d, s, __fmul_rn ( fBase, tLeaf.m_fIDF ), uFields, 0u );
						}
						bDirty = true;
					} else
					{
						if ( tOp.m_eCode==OP_TERM_SET )
						{
							#pragma unroll
							for ( int k=0; k<TILE_W/EVAL_THREADS; ++k )
								V.Cnt ( d, k*EVAL_THREADS+tid ) = 0;	// slot-local
							bDirty = true;
						}
						if ( bDirty )
							__syncthreads();
						const uint32_t uEmitBit = 1u<<tOp.m_uLeaf;
						if ( ( q.m_uPreMask>>tOp.m_uLeaf ) & 1u )
						{
							// scatter the predecoded entries of this keyword, one entry per thread
							const int iFrom = S.m_dPreOff[tOp.m_uLeaf]*32, iTo = iFrom + (int)( b1-b0 )*32;
							for ( int i=iFrom+tid; i<iTo; i+=EVAL_THREADS )
							{
								const PreEntry_t tEntry = pPre[i];
								if ( tEntry.m_uRowid==0xFFFFFFFFu )
									continue;
								const int s = (int)( tEntry.m_uRowid-uTileLo );
								const bool bOn = ApplyTermOp<HITS> ( V, tOp, d, s, tEntry.m_fTf, tEntry.m_uFields, uEmitBit );
								if ( HITS && bOn )
								{
									pHitpos[(size_t)tOp.m_uLeaf*TILE_W+s] = pPreHitpos[i];
									pLeafTf[(size_t)tOp.m_uLeaf*TILE_W+s] = (uint64_t)__float_as_uint ( tEntry.m_fTf ) | ( (uint64_t)tEntry.m_uFields<<32 );
								}
							}
						} else
						{
							const bool bSkippable = bChain && tOp.m_eCode==OP_TERM_AND;
							for ( uint32_t b=b0+iWarp; b<b1; b+=EVAL_WARPS )
							{
								if ( bSkippable )
								{
									// no candidate inside this block's rowid range [base_b, base_b+1): never decode it
									const uint32_t uLo = max ( __ldg ( tIdx.m_pBlkRowid+tLeaf.m_uFirstBlk+b ), uTileLo ) - uTileLo;
									const uint32_t uHi = ( b+1<tLeaf.m_nBlocks ? min ( __ldg ( tIdx.m_pBlkRowid+tLeaf.m_uFirstBlk+b+1 ), uTileHi ) : uTileHi ) - uTileLo;
									if ( !AnyAliveInRange ( S.m_dAliveBits, uLo, uHi, iLane ) )
										continue;
								}
								DecodedDoc_t tDoc;
								DecodeBlock<HITS> ( tIdx, tLeaf, b, S.m_dStage[iWarp], S.m_dRecStart[iWarp], iLane, tDoc );
								const uint32_t uFields = tDoc.m_uFields & tLeaf.m_uQueriedFields;
								if ( !tDoc.m_bValid || tDoc.m_uRowid<uTileLo || tDoc.m_uRowid>=uTileHi || !uFields )
									continue;
								if ( HITS && tLeaf.m_iTermPos && !HasAcceptableHit ( tIdx.m_pSpp, tDoc.m_uHitlistPos, tLeaf.m_uQueriedFields, tLeaf.m_iTermPos ) )
									continue;
								const int s = (int)( tDoc.m_uRowid-uTileLo );
								// ExtTerm_T::GetDocsChunk, src/searchnode.cpp:1946
								const float fHits = __uint2float_rn ( tDoc.m_uHits );
								const float fTf = __fmul_rn ( __fdiv_rn ( fHits, __fadd_rn ( fHits, 1.2f ) ), tLeaf.m_fIDF );
								const bool bOn = ApplyTermOp<HITS> ( V, tOp, d, s, fTf, uFields, uEmitBit );
								if ( HITS && bOn )
								{
									pHitpos[(size_t)tOp.m_uLeaf*TILE_W+s] = tDoc.m_uHitlistPos;
									pLeafTf[(size_t)tOp.m_uLeaf*TILE_W+s] = (uint64_t)__float_as_uint ( fTf ) | ( (uint64_t)uFields<<32 );
								}
							}
						}
						__syncthreads();
						bDirty = false;
					}
				} else if ( tOp.m_eCode==OP_NWAY )
				{
					// ExtNWay_T::GetDocsChunk, src/searchnode.cpp:3805-3848: candidates = AND of the node's keywords (already in v[dst]);
					// a candidate survives iff the acceptor emits at least one folded hit
					if constexpr ( HITS )
					{
						const int j = tOp.m_uArg;
						const DevNWay_t & tN = q.m_dNWay[j];
						const int n = CompactAlive ( V, d, tOp.m_uAliveDst, TILE_W, &S.m_iListCnt );
						for ( int i=tid; i<n; i+=EVAL_THREADS )
						{
							const int s = V.m_pList[i];
							const uint32_t uOn = V.Emit ( d, s );	// the keywords sitting on the document
							bool bOk = true;
							if ( tN.m_eKind>=NWAY_NOTNEAR )
							{
								// optional children (NOTNEAR's right side, quorum keywords): mark the absent ones; quorum: count
								// (ExtQuorum_c: a keyword the query repeats counts as often as it is repeated, up to its hits on the document, :4602-4632)
								int iQuorum = 0;
								for ( int w=( tN.m_eKind==NWAY_NOTNEAR ? 1 : 0 ); w<tN.m_nWords; ++w )
								{
									const int l = tN.m_dLeaf[w];
									if ( !( ( uOn>>l ) & 1u ) )
									{
										pHitpos[(size_t)l*TILE_W+s] = HITPOS_ABSENT;
										continue;
									}
									int k = 1;
									if ( tN.m_dCount[w]>1 )
									{
										HitCursor_t c;
										SeekHitlist ( c, tIdx.m_pSpp, pHitpos[(size_t)l*TILE_W+s] );
										k = 0;
										while ( k<(int)tN.m_dCount[w] && NextHit ( c, q.m_dLeaves[l].m_uQueriedFields, q.m_dLeaves[l].m_iTermPos ) )
											++k;
									}
									iQuorum += k;
								}
								if ( tN.m_eKind==NWAY_QUORUM )
									bOk = iQuorum>=max ( tN.m_iOpArg, 1 );
							}
							DocHits_t H;
							if ( bOk )
							{
								NWayOpen ( tIdx, q, j, pHitpos, TILE_W, s, H );
								bOk = H.m_dNWay[j].m_tHead.m_uHitpos!=0;
							}
							if ( bOk )
							{
								if ( tN.m_eKind==NWAY_BEFORE || tN.m_eKind==NWAY_NOTNEAR )
								{
									// the document carries its first child's TF*IDF and fields only (ExtOrder_c :4914, ExtNotNear_c :5430)
									const uint64_t uRec = pLeafTf[(size_t)tN.m_dLeaf[0]*TILE_W+s];
									V.Tfidf ( d, s ) = __uint_as_float ( (uint32_t)uRec );
									V.Fields ( d, s ) = (uint32_t)( uRec>>32 );
								} else if ( tN.m_eKind==NWAY_QUORUM )
								{
									// TF*IDF is summed in the order of the reference's children vector, which RemoveFast reshuffles whenever a
									// keyword runs out of documents (ExtQuorum_c::GetDocsChunk :4486-4537): the order of this rowid's interval
									const uint32_t uRow = uTileLo+(uint32_t)s;
									int iInt = 0;
									while ( iInt+1<S.m_nQIntervals && uRow>S.m_dQBound[iInt] )
										++iInt;
									float fT = 0.0f;
									uint32_t uF = 0;
									bool bFirst = true;
									for ( int k=0; k<S.m_dQLen[iInt]; ++k )
									{
										const int l = tN.m_dLeaf[S.m_dQOrder[iInt][k]];
										if ( !( ( uOn>>l ) & 1u ) )
											continue;
										const uint64_t uRec = pLeafTf[(size_t)l*TILE_W+s];
										const float fTf = __uint_as_float ( (uint32_t)uRec );
										fT = bFirst ? fTf : __fadd_rn ( fT, fTf );
										uF |= (uint32_t)( uRec>>32 );
										bFirst = false;
									}
									V.Tfidf ( d, s ) = fT;
									V.Fields ( d, s ) = uF;
								} else
									V.Fields ( d, s ) = 1u<<( ( H.m_dNWay[j].m_uFirstRawHit>>24 ) & 31u );
								V.Emit ( d, s ) = 1u<<( EMIT_NWAY_SHIFT+j );
								V.Cnt ( d, s ) = tOp.m_uAliveOut;
							} else
								V.Cnt ( d, s ) = 0;
						}
						__syncthreads();
						bDirty = false;
					}
				} else
				{
					const int r = tOp.m_uSrc;
					const uint8_t uAd = tOp.m_uAliveDst, uAs = tOp.m_uAliveSrc;
					#pragma unroll 2
					for ( int s=tid; s<TILE_W; s+=EVAL_THREADS )
					{
						const bool bD = V.Cnt ( d, s )==uAd, bS = V.Cnt ( r, s )==uAs;
						switch ( tOp.m_eCode )
						{
						case OP_VEC_AND:
							if ( bD && bS )
							{
								V.Tfidf ( d, s ) = __fadd_rn ( V.Tfidf ( d, s ), V.Tfidf ( r, s ) );
								V.Fields ( d, s ) |= V.Fields ( r, s );
								if ( HITS ) V.Emit ( d, s ) |= V.Emit ( r, s );
							} else if ( bD )
								V.Cnt ( d, s ) = 0;
							break;
						case OP_VEC_OR:
							if ( bD && bS )
							{
								V.Tfidf ( d, s ) = __fadd_rn ( V.Tfidf ( d, s ), V.Tfidf ( r, s ) );
								V.Fields ( d, s ) |= V.Fields ( r, s );
								if ( HITS ) V.Emit ( d, s ) |= V.Emit ( r, s );
							} else if ( bS )
							{
								V.Tfidf ( d, s ) = V.Tfidf ( r, s ); V.Fields ( d, s ) = V.Fields ( r, s ); V.Cnt ( d, s ) = uAd;
								if ( HITS ) V.Emit ( d, s ) = V.Emit ( r, s );
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
								if ( HITS ) V.Emit ( d, s ) |= V.Emit ( r, s );
							}
							break;
						default:
							break;
						}
					}
					bDirty = true;
				}
			}

			// where to go next (read the shared cursor state now: the barrier at the end of the tile separates these reads
			// from the next tile's writes)
			uint32_t uNextLo = uTileLo+TILE_W;
			if ( bJump )
			{
				const DevLeaf_t & tDrv = q.m_dLeaves[q.m_iDriverLeaf];
				const uint32_t b1 = S.m_dLeafB1[q.m_iDriverLeaf];
				uint32_t uNext = S.m_uNextRow;
				if ( b1<tDrv.m_nBlocks )
					uNext = min ( uNext, __ldg ( tIdx.m_pBlkRowid+tDrv.m_uFirstBlk+b1 ) );
				uNextLo = uNext==0xFFFFFFFFu ? 0xFFFFFFFFu : max ( uNextLo, uNext & ~(uint32_t)( TILE_W-1 ) );
			}

			// rank + filter + push survivors of this tile
			{
				const uint8_t uAlive = (uint8_t)q.m_uAliveRoot;
				const Key128_t tThr = S.m_tThr;
				Key128_t * pPool = pPool0 + (size_t)S.m_iPoolBuf*P.m_iPoolCap;
				const int nSlots = (int)( uTileHi-uTileLo );
				// hit-consuming rankers walk a compacted candidate list (one document per thread); the others scan the slots
				const bool bStateRanker = HITS && q.m_bStateRanker;
				int nWork = nSlots;
				if ( bStateRanker )
					nWork = CompactAlive ( V, 0, uAlive, nSlots, &S.m_iListCnt );
				for ( int sBase=0; sBase<nWork; sBase+=EVAL_THREADS )
				{
					int s = sBase+tid;
					bool bPush = false;
					Key128_t tKey;
					bool bCand = s<nWork;
					if ( bCand )
					{
						if ( bStateRanker )
							s = V.m_pList[s];
						else
							bCand = V.Cnt ( 0, s )==uAlive;
					}
					if ( bCand )
					{
						const uint32_t uRowid = uTileLo+s;
						bool bOk = PassFilters ( tIdx, q, uRowid );
						int iWeight = 1;	// ExtRanker_None_c, src/sphinxsearch.cpp:1160
						if ( bOk && q.m_eRanker!=2 )
						{
							// seed weight src/sphinxsearch.cpp:1070
							const int iSeed = __float2int_rz ( __fmul_rn ( __fadd_rn ( V.Tfidf ( 0, s ), 0.5f ), 1000.0f ) );
							if ( bStateRanker )
							{
								if constexpr ( HITS )
								{
									DocHits_t H;
									bOk = RankDocByHits ( tIdx, q, V.Emit ( 0, s ), pHitpos, TILE_W, s, iSeed, H, iWeight, uMyHitBytes );
								}
							} else
							{
								// ExtRanker_WeightSum_c :1112-1129
								uint32_t uMask = V.Fields ( 0, s );
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
			}
			__syncthreads();

			uTileLo = uNextLo;
		}

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
			if ( HITS && P.m_pWork )
			{
				#pragma unroll
				for ( int d=16; d; d>>=1 )
					uMyHitBytes += __shfl_xor_sync ( FULL_MASK, uMyHitBytes, d );
				if ( iLane==0 && uMyHitBytes )
					atomicAdd ( P.m_pWork, uMyHitBytes );
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

#include "stream_kernel.cuh"
#include "orbits_kernel.cuh"

//////////////////////////////////////////////////////////////////////////
// K2: driver-led intersection kernel for pure AND queries (ExtMultiAnd_T::AdvanceQwords, src/searchnode.cpp:2864-2889;
// the skiplist jump of DiskIndexQword_c::HintRowID/AdvanceTo, src/sphinx.cpp:391-451)
//////////////////////////////////////////////////////////////////////////

static const int AND_CHUNK_BLOCKS = 128;	///< driver blocks per CTA round: 4096 candidates = what the candidate pool can take

struct AndShared_t
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
	uint32_t		m_dRows[EVAL_WARPS][32];	///< the other keyword's decoded block, for the candidates' lookups
	uint32_t		m_dHits[EVAL_WARPS][32];
	uint32_t		m_dFields[EVAL_WARPS][32];
	uint64_t		m_dHitposSm[EVAL_WARPS][32];	///< hit variant: hitlist positions of the probed block
	uint16_t		m_dRecStart[EVAL_WARPS][34];
	__align__(16) uint8_t m_dStage[EVAL_WARPS][STAGE_BYTES];
};

/// Does the keyword hold the lane's candidate row? Warp-cooperative: hot keywords are probed in the dense store; sparse ones by a
/// binary search of the resident block table per candidate (FindSpan, src/sphinx.cpp:407-451), then only the distinct blocks that
/// may hold a candidate are decoded (each once per warp) and searched. Returns the queried field mask (0 = absent) and the hit count.
/// per-thread hit machinery only exists in the hit variant
template<bool HITS> struct HitsState_T { struct Type_t {}; };
template<> struct HitsState_T<true> { typedef DocHits_t Type_t; };

template<bool HITS>
__device__ __noinline__ uint32_t WarpProbeKeyword ( const DevIndex_t & tIdx, const DevHotStore_t & tHot, const DevLeaf_t & tLeaf, uint32_t uRowid, bool bActive,
	AndShared_t & S, int iWarp, int iLane, uint32_t & uHitsOut, uint64_t & uHitposOut, uint32_t uPreloaded=0xFFFFFFFFu )
{
	uint32_t uHits = 0, uF = 0;
	uint64_t uHitpos = 0;
	if ( !HITS && tLeaf.m_iHot>=0 )
	{
		if ( bActive )
		{
			// (uPreloaded: the row's store entry, requested by ProbeGroup before the group's first probe)
			const uint32_t v = uPreloaded!=0xFFFFFFFFu ? uPreloaded : __ldg ( tHot.m_pData + (size_t)tLeaf.m_iHot*tHot.m_iStride + uRowid );
			uHits = v & 255u;
			uF = uHits ? ( ( v>>8 ) & tLeaf.m_uQueriedFields ) : 0u;
			if ( uF && uHits==255u )
				uHits = HotEscapeHits ( tHot, tLeaf.m_iHot, uRowid );
		}
	} else if ( tLeaf.m_nBlocks )
	{
		const uint32_t * pBase = tIdx.m_pBlkRowid + tLeaf.m_uFirstBlk;
		uint32_t uBlk = 0;
		if ( bActive )
		{
			uint32_t lo = 0, hi = tLeaf.m_nBlocks;	// first index with base > rowid
			while ( lo<hi )
			{
				const uint32_t mid = lo + ( ( hi-lo )>>1 );
				if ( __ldg ( pBase+mid )>uRowid ) hi = mid; else lo = mid+1;
			}
			uBlk = lo-1;	// base of block 0 is 0, so lo>=1
		}
		unsigned uTodo = __ballot_sync ( FULL_MASK, bActive );
		while ( uTodo )
		{
			const int iLeader = __ffs ( uTodo )-1;
			const uint32_t uSel = __shfl_sync ( FULL_MASK, uBlk, iLeader );
			DecodedDoc_t tOther;
			DecodeBlock<HITS> ( tIdx, tLeaf, uSel, S.m_dStage[iWarp], S.m_dRecStart[iWarp], iLane, tOther );
			S.m_dRows[iWarp][iLane] = tOther.m_bValid ? tOther.m_uRowid : 0xFFFFFFFFu;
			S.m_dHits[iWarp][iLane] = tOther.m_uHits;
			S.m_dFields[iWarp][iLane] = tOther.m_uFields;
			if ( HITS )
				S.m_dHitposSm[iWarp][iLane] = tOther.m_uHitlistPos;
			__syncwarp();
			const bool bMine = bActive && uBlk==uSel;
			if ( bMine )
			{
				// rows ascend (invalid tail = 0xFFFFFFFF): lower bound over 32 entries
				int l = 0;
				#pragma unroll
				for ( int iStep=16; iStep; iStep>>=1 )
					if ( S.m_dRows[iWarp][l+iStep-1]<uRowid )
						l += iStep;
				if ( S.m_dRows[iWarp][l]==uRowid )
				{
					uHits = S.m_dHits[iWarp][l];
					uF = S.m_dFields[iWarp][l] & tLeaf.m_uQueriedFields;
					if ( HITS )
						uHitpos = S.m_dHitposSm[iWarp][l];
				}
			}
			uTodo &= ~__ballot_sync ( FULL_MASK, bMine );
			__syncwarp();
		}
	}
	uHitsOut = uHits;
	uHitposOut = uHitpos;
	return uF;
}

/// one AND group (ops [iOp0, iOp0+nOps) = SET, AND, AND...) evaluated on the lanes' candidate rows by probing, in the reference's
/// rarest-first order; iSkipOp = op whose keyword is already accounted for (the driver), -1 = probe all. Accumulates into fT / uFields.
template<bool HITS>
__device__ __forceinline__ bool ProbeGroup ( const DevIndex_t & tIdx, const DevHotStore_t & tHot, const DevQuery_t & q, int iOp0, int nOps, int iSkipOp,
	uint32_t uRowid, bool bActive, AndShared_t & S, int iWarp, int iLane, float & fT, uint32_t & uFields, uint64_t * pLaneHitpos )
{
	for ( int iOp=iOp0; iOp<iOp0+nOps; ++iOp )
	{
		if ( iOp==iSkipOp )
			continue;
		if ( !__any_sync ( FULL_MASK, bActive ) )
			break;
		const DevLeaf_t & tLeaf = q.m_dLeaves[q.m_dOps[iOp].m_uLeaf];
		uint32_t uHits;
		uint64_t uHitpos;
		const uint32_t uF = WarpProbeKeyword<HITS> ( tIdx, tHot, tLeaf, uRowid, bActive, S, iWarp, iLane, uHits, uHitpos );
		if ( q.m_dOps[iOp].m_eCode==OP_TERM_ANDNOT )
		{
			// `a b -c`: ExtAndNot_c keeps the left side's documents that the right side does not hold, with the left side's weight and fields
			// (src/searchnode.cpp:3618-3706)
			bActive = bActive && !uF;
			continue;
		}
		if ( bActive )
		{
			if ( uF )
			{
				if ( HITS )
					pLaneHitpos[(size_t)q.m_dOps[iOp].m_uLeaf*32] = uHitpos;
				// ExtMultiAnd_T::GetDocsChunk, src/searchnode.cpp:2821-2832
				const float fHits = __uint2float_rn ( uHits );
				const float fTf = __fmul_rn ( uHits<255 ? S.m_dTf[uHits & 255u] : __fdiv_rn ( fHits, __fadd_rn ( fHits, 1.2f ) ), tLeaf.m_fIDF );
				fT = ( iOp==iOp0 ) ? fTf : __fadd_rn ( fT, fTf );
				uFields |= uF;
			} else
				bActive = false;
		}
	}
	return bActive;
}

/// Work item = (query, range of the rarest keyword's 32-doc blocks). A warp decodes one driver block: 32 candidate rows, one
/// per lane, kept in registers. For every other keyword, in the reference's rarest-first order: hot keywords are probed in the
/// dense store; sparse ones by a binary search of the resident skiplist (block table) per candidate, then only the blocks that
/// hold a candidate are decoded (each once per warp) and searched. TF*IDF accumulates in that same order.
template<bool HITS>
__global__ void __launch_bounds__ ( EVAL_THREADS, 3 ) and_kernel ( EvalParams_t P )
{
	__shared__ AndShared_t S;
	const int tid = threadIdx.x, iWarp = tid>>5, iLane = tid & 31;
	const DevIndex_t & tIdx = P.m_tIndex;
	Key128_t * pPool0 = P.m_pPool + (size_t)blockIdx.x*2*P.m_iPoolCap;
	// hit variant: hitlist position of (keyword, candidate lane), [MAX_LEAVES][32] per warp, in the CTA's hit scratch
	uint64_t * pLaneHitpos = HITS ? P.m_pHitpos + ( (size_t)blockIdx.x*EVAL_WARPS+iWarp )*MAX_LEAVES*32 + iLane : nullptr;
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
		}
		int iMyTotal = 0;
		unsigned long long uMyHitBytes = 0;
		// the item's AND group (DNF programs: OR of AND groups; a pure AND query is one group) and its driver = rarest keyword
		const int iGroup = (int)tItem.m_uPad;
		const int iOp0 = q.m_dGroupOp0[iGroup], nGroupOps = q.m_dGroupOps[iGroup];
		const DevLeaf_t & tDrv = q.m_dLeaves[q.m_dOps[iOp0].m_uLeaf];

		for ( uint32_t uChunk=tItem.m_uRowLo; uChunk<tItem.m_uRowHi; uChunk+=AND_CHUNK_BLOCKS )
		{
			// compact the candidate pool if this round could overflow it
			__syncthreads();
			const int iPoolNow = S.m_iPoolCnt;
			__syncthreads();	// nobody pushes before everybody has read the level
			if ( iPoolNow+AND_CHUNK_BLOCKS*32>P.m_iPoolCap )
			{
				Key128_t * pIn = pPool0 + (size_t)S.m_iPoolBuf*P.m_iPoolCap;
				Key128_t * pOut = pPool0 + (size_t)( S.m_iPoolBuf^1 )*P.m_iPoolCap;
				Key128_t tThr = CtaSelectTopK ( pIn, iPoolNow, iK, pOut, S.m_tSel );
				if ( tid==0 )
				{
					S.m_tThr = tThr;
					S.m_iPoolCnt = iK;
					S.m_iPoolBuf ^= 1;
				}
				__syncthreads();
			}
			const Key128_t tThr = S.m_tThr;
			Key128_t * pPool = pPool0 + (size_t)S.m_iPoolBuf*P.m_iPoolCap;
			const uint32_t uChunkEnd = min ( uChunk+(uint32_t)AND_CHUNK_BLOCKS, tItem.m_uRowHi );

			for ( uint32_t b=uChunk+iWarp; b<uChunkEnd; b+=EVAL_WARPS )
			{
				// the driver block: 32 candidates
				typename HitsState_T<HITS>::Type_t tHits;
				DecodedDoc_t tDoc;
				DecodeBlock<HITS> ( tIdx, tDrv, b, S.m_dStage[iWarp], S.m_dRecStart[iWarp], iLane, tDoc );
				const uint32_t uRowid = tDoc.m_uRowid;
				uint32_t uFields = tDoc.m_uFields & tDrv.m_uQueriedFields;
				bool bAlive = tDoc.m_bValid && uFields;
				if ( HITS )
					pLaneHitpos[(size_t)q.m_dOps[iOp0].m_uLeaf*32] = tDoc.m_uHitlistPos;
				float fTfidf;
				{
					const float fHits = __uint2float_rn ( tDoc.m_uHits );
					fTfidf = __fmul_rn ( tDoc.m_uHits<255 ? S.m_dTf[tDoc.m_uHits & 255u] : __fdiv_rn ( fHits, __fadd_rn ( fHits, 1.2f ) ), tDrv.m_fIDF );
				}

				// the rest of this group
				bAlive = ProbeGroup<HITS> ( tIdx, P.m_tHot, q, iOp0, nGroupOps, iOp0, uRowid, bAlive, S, iWarp, iLane, fTfidf, uFields, pLaneHitpos );

				// DNF: a document that also matches an earlier group is emitted by that group's items; later groups that match add
				// their TF*IDF in group order (ExtOr_c: left + right, src/searchnode.cpp:3486-3504)
				for ( int g=0; g<q.m_nGroups && q.m_nGroups>1; ++g )
				{
					if ( g==iGroup )
						continue;
					if ( !__any_sync ( FULL_MASK, bAlive ) )
						break;
					float fT = 0.0f;
					uint32_t uF = 0;
					const bool bMatch = ProbeGroup<false> ( tIdx, P.m_tHot, q, q.m_dGroupOp0[g], q.m_dGroupOps[g], -1, uRowid, bAlive, S, iWarp, iLane, fT, uF, nullptr );
					if ( bMatch )
					{
						if ( g<iGroup )
							bAlive = false;
						else
						{
							fTfidf = __fadd_rn ( fTfidf, fT );
							uFields |= uF;
						}
					}
				}

				// hit variant: the phrase / proximity acceptor over the chain's keywords (ExtNWay_T::GetDocsChunk, src/searchnode.cpp:3805-3848),
				// then ranking by the document's hit stream; one candidate document per lane
				uint32_t uEmit = 0;
				if constexpr ( HITS )
				{
					for ( int iOp=iOp0; iOp<iOp0+nGroupOps; ++iOp )
						uEmit |= 1u<<q.m_dOps[iOp].m_uLeaf;
					if ( q.m_nNWay>0 && bAlive )
					{
						const int j = q.m_dOps[iOp0+nGroupOps].m_uArg;
						NWayOpen ( tIdx, q, j, pLaneHitpos-iLane, 32, iLane, tHits );
						if ( tHits.m_dNWay[j].m_tHead.m_uHitpos )
						{
							uFields = 1u<<( ( tHits.m_dNWay[j].m_uFirstRawHit>>24 ) & 31u );
							uEmit = 1u<<( EMIT_NWAY_SHIFT+j );
						} else
							bAlive = false;
					}
				}

				// rank + filter + push
				bool bPush = false;
				Key128_t tKey;
				if ( bAlive )
				{
					bool bOk = PassFilters ( tIdx, q, uRowid );
					int iWeight = 1;
					if ( bOk && q.m_eRanker!=2 )
					{
						const int iSeed = __float2int_rz ( __fmul_rn ( __fadd_rn ( fTfidf, 0.5f ), 1000.0f ) );
						if ( HITS && q.m_bStateRanker )
						{
							if constexpr ( HITS )
								bOk = RankDocByHits ( tIdx, q, uEmit, pLaneHitpos-iLane, 32, iLane, iSeed, tHits, iWeight, uMyHitBytes );
						} else
						{
							uint32_t uRank = 0;
							if ( !uFields )
								uRank = 1;
							else if ( q.m_nWeights<=4 )
								uRank = S.m_dRankTab[uFields & 15u];
							else
								for ( int i=0; i<q.m_nWeights; ++i )
									if ( uFields & ( 1u<<i ) )
										uRank += (uint32_t)q.m_dWeights[i];
							iWeight = q.m_eRanker==4 ? (int)uRank : (int)( (uint32_t)iSeed + uRank*1000u );	// ExtRanker_WeightSum_c<false> under SPH_RANK_PROXIMITY
						}
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
			if ( HITS && P.m_pWork )
			{
				#pragma unroll
				for ( int d=16; d; d>>=1 )
					uMyHitBytes += __shfl_xor_sync ( FULL_MASK, uMyHitBytes, d );
				if ( iLane==0 && uMyHitBytes )
					atomicAdd ( P.m_pWork, uMyHitBytes );
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

//////////////////////////////////////////////////////////////////////////
// per-query merge of the work items' candidates: select K, sort best first
//////////////////////////////////////////////////////////////////////////

__device__ void CtaBitonicSortDesc ( Key128_t * p, int n /* power of two */ )
{
	const int tid = threadIdx.x, nThreads = blockDim.x;
	for ( int k=2; k<=n; k<<=1 )
		for ( int j=k>>1; j>0; j>>=1 )
		{
			for ( int i=tid; i<n; i+=nThreads )
			{
				int l = i ^ j;
				if ( l>i )
				{
					Key128_t a = p[i], b = p[l];
					bool bDesc = ( i & k )==0;
					if ( bDesc ? KeyLess ( a, b ) : KeyLess ( b, a ) )
					{
						p[i] = b; p[l] = a;
					}
				}
			}
			__syncthreads();
		}
}

__global__ void __launch_bounds__ ( 256 ) merge_kernel ( MergeParams_t P )
{
	__shared__ SelectSmem_t tSel;
	__shared__ int iTotalKeys;
	__shared__ unsigned long long uTotalFound;
	const int tid = threadIdx.x;

	for ( int iQuery=blockIdx.x; iQuery<P.m_nQueries; iQuery+=gridDim.x )
	{
		const DevQueryCore_t & q = P.m_pQueries[iQuery];
		const int iK = q.m_iMaxMatches;
		Key128_t * pA = P.m_pScratch + (size_t)iQuery*P.m_iScratchStride;	// [0, stride/2) gather, [stride/2, stride) select output
		const int iHalf = P.m_iScratchStride/2;
		Key128_t * pB = pA+iHalf;

		if ( tid==0 )
		{
			iTotalKeys = 0;
			uTotalFound = 0;
		}
		__syncthreads();
		// gather
		for ( int it=0; it<q.m_nItems; ++it )
		{
			const int iItem = q.m_iFirstItem+it;
			const int n = P.m_pItemOut[iItem].m_nKeys;
			const int iBase = iTotalKeys;
			const Key128_t * pSrc = P.m_pItemKeys + (size_t)iItem*P.m_iKMax;
			for ( int i=tid; i<n; i+=blockDim.x )
				pA[iBase+i] = pSrc[i];
			__syncthreads();
			if ( tid==0 )
			{
				iTotalKeys = iBase+n;
				uTotalFound += (unsigned long long)P.m_pItemOut[iItem].m_iTotalFound;
			}
			__syncthreads();
		}
		int n = iTotalKeys;
		Key128_t * pCur = pA;
		if ( n>iK )
		{
			CtaSelectTopK ( pA, n, iK, pB, tSel );
			n = iK;
			pCur = pB;
		}
		int nPad = 1;
		while ( nPad<n )
			nPad <<= 1;
		for ( int i=n+tid; i<nPad; i+=blockDim.x )
		{
			pCur[i].m_uHi = 0; pCur[i].m_uLo = 0;
		}
		__syncthreads();
		CtaBitonicSortDesc ( pCur, nPad );

		const int iSlot = P.m_pOutSlot[iQuery];
		Key128_t * pOut = P.m_pOutKeys + (size_t)iSlot*P.m_iKMax;
		int64_t * pDocid = P.m_pOutDocid + (size_t)iSlot*P.m_iKMax;
		for ( int i=tid; i<n; i+=blockDim.x )
		{
			Key128_t k = pCur[i];
			pOut[i] = k;
			uint32_t uRowid = ~(uint32_t)( k.m_uLo>>32 ) - P.m_tIndex.m_uRowidBase;
			const uint32_t * pRow = P.m_tIndex.m_pSpa + (size_t)uRowid*P.m_tIndex.m_iStride;
			pDocid[i] = (int64_t)( (uint64_t)pRow[0] | ( (uint64_t)pRow[1]<<32 ) );	// `id` is always attribute 0, bigint
		}
		if ( tid==0 )
		{
			P.m_pOutCount[iSlot] = n;
			P.m_pOutTotal[iSlot] = (int64_t)uTotalFound;
		}
		__syncthreads();
	}
}

//////////////////////////////////////////////////////////////////////////
// K9: merge of per-shard top-K key lists (disjoint rowid ranges => no dedupe, src/searchd.cpp:3910-3952)
//////////////////////////////////////////////////////////////////////////

__global__ void __launch_bounds__ ( 256 ) shard_merge_kernel ( const Key128_t * pKeys, const int32_t * pCounts, int nShards, int nQueries, int iK,
	Key128_t * pScratch, int iScratchStride, Key128_t * pOutKeys, int32_t * pOutCounts,
	const int64_t * pDocids, const uint32_t * pShardBase, int64_t * pOutDocid )
{
	__shared__ SelectSmem_t tSel;
	__shared__ int iTotalKeys;
	const int tid = threadIdx.x;
	for ( int iQuery=blockIdx.x; iQuery<nQueries; iQuery+=gridDim.x )
	{
		Key128_t * pA = pScratch + (size_t)iQuery*iScratchStride;
		Key128_t * pB = pA + iScratchStride/2;
		if ( tid==0 )
			iTotalKeys = 0;
		__syncthreads();
		for ( int s=0; s<nShards; ++s )
		{
			const int n = min ( pCounts[(size_t)s*nQueries+iQuery], iK );
			const int iBase = iTotalKeys;
			const Key128_t * pSrc = pKeys + ( (size_t)s*nQueries+iQuery )*iK;
			for ( int i=tid; i<n; i+=blockDim.x )
				pA[iBase+i] = pSrc[i];
			__syncthreads();
			if ( tid==0 )
				iTotalKeys = iBase+n;
			__syncthreads();
		}
		int n = iTotalKeys;
		Key128_t * pCur = pA;
		if ( n>iK )
		{
			CtaSelectTopK ( pA, n, iK, pB, tSel );
			n = iK;
			pCur = pB;
		}
		int nPad = 1;
		while ( nPad<n )
			nPad <<= 1;
		for ( int i=n+tid; i<nPad; i+=blockDim.x )
		{
			pCur[i].m_uHi = 0; pCur[i].m_uLo = 0;
		}
		__syncthreads();
		CtaBitonicSortDesc ( pCur, nPad );
		for ( int i=tid; i<n; i+=blockDim.x )
		{
			const Key128_t tKey = pCur[i];
			pOutKeys[(size_t)iQuery*iK+i] = tKey;
			if ( pDocids )
			{
				// the match's document id travels with its shard's list: the owner follows from the global rowid, the entry from a
				// binary search of that (best-first) list
				const uint32_t uRow = ~(uint32_t)( tKey.m_uLo>>32 );
				int s = 0;
				while ( s+1<nShards && uRow>=pShardBase[s+1] )
					++s;
				const Key128_t * pSrc = pKeys + ( (size_t)s*nQueries+iQuery )*iK;
				int lo = 0, hi = min ( pCounts[(size_t)s*nQueries+iQuery], iK );
				while ( lo<hi )
				{
					const int mid = ( lo+hi )>>1;
					if ( KeyLess ( tKey, pSrc[mid] ) ) lo = mid+1; else hi = mid;
				}
				pOutDocid[(size_t)iQuery*iK+i] = pDocids [ ( (size_t)s*nQueries+iQuery )*iK + lo ];
			}
		}
		if ( tid==0 )
			pOutCounts[iQuery] = n;
		__syncthreads();
	}
}

//////////////////////////////////////////////////////////////////////////
// host-callable launchers (engine.cpp is plain C++ and never sees <<< >>>)
//////////////////////////////////////////////////////////////////////////

/// the CTA-per-tile evaluator is only instantiated for hit-consuming queries; doc-only ones run on stream_kernel / and_kernel
size_t EvalDynSmemBytes ( int nStack )
{
	return (size_t)nStack*TILE_W*13 + (size_t)TILE_W*2;
}

cudaError_t LaunchEval ( const EvalParams_t & P, int nStack, int nCtas, cudaStream_t tStream )
{
	size_t iDyn = EvalDynSmemBytes ( nStack );
	cudaError_t e = cudaFuncSetAttribute ( eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iDyn );
	if ( e!=cudaSuccess )
		return e;
	eval_kernel<true><<<nCtas, EVAL_THREADS, iDyn, tStream>>> ( P, nStack );
	return cudaGetLastError();
}

int EvalOccupancy ( int nStack )
{
	int n = 0;
	size_t iDyn = EvalDynSmemBytes ( nStack );
	cudaFuncSetAttribute ( eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iDyn );
	if ( cudaOccupancyMaxActiveBlocksPerMultiprocessor ( &n, eval_kernel<true>, EVAL_THREADS, iDyn )!=cudaSuccess )
		return 1;
	return n>0 ? n : 1;
}

/// single-level programs (pure OR / single keyword) take 512-row mini-tiles, deeper ones 256-row (shared memory per level)
static int StreamMiniWidth ( int nStack )
{
	return nStack<=1 ? 512 : 256;
}

/// the register-OR class needs per warp: sparse overlay 512*4 + candidate rows 256 + queue (32+256)*4
static const size_t OR_WARP_SMEM = 512*4 + 256 + 288*4;

/// sparse postings one mini-tile of the bound + exact pass kernels can hold, per warp (sizes EvalParams_t::m_pOrList)
int StreamOrListCap ( int iMode )
{
	return iMode==3 ? 0 : OR_LIST_CAP;	// (orbits_kernel reads the decoded posting lists)
}

size_t StreamDynSmemBytes ( int nStack, int iMode )
{
	if ( iMode==3 )
		return (size_t)EVAL_WARPS*OB_WARP_SMEM;
	if ( iMode )
		return (size_t)EVAL_WARPS*512*9;	// (the kernel strides its warps by nStack*MINI_W*9 = 4608 >= OR_WARP_SMEM)
	return (size_t)nStack*EVAL_WARPS*StreamMiniWidth ( nStack )*9;
}

template<typename KERNEL>
static cudaError_t LaunchStreamT ( KERNEL fnKernel, const EvalParams_t & P, int nStack, size_t iDyn, int nCtas, cudaStream_t tStream )
{
	cudaError_t e = cudaFuncSetAttribute ( fnKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iDyn );
	if ( e!=cudaSuccess )
		return e;
	fnKernel<<<nCtas, EVAL_THREADS, iDyn, tStream>>> ( P, nStack );
	return cudaGetLastError();
}

/// iMode: 0 = general tile program, 1 = pure OR programs, 2 = DNF programs with hot multi-keyword groups (bound + exact pass),
/// 3 = pure OR programs over the presence bitmaps (orbits_kernel)
cudaError_t LaunchStream ( const EvalParams_t & P, int nStack, int iMode, int nCtas, cudaStream_t tStream )
{
	static_assert ( OR_WARP_SMEM<=512*9, "register-OR scratch must fit the warp's slice" );
	const size_t iDyn = StreamDynSmemBytes ( nStack, iMode );
	if ( iMode==3 )
	{
		auto fnKernel = P.m_tHot.m_nBitFields<=2 ? orbits_kernel<2> : orbits_kernel<4>;
		cudaError_t e = cudaFuncSetAttribute ( fnKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iDyn );
		if ( e!=cudaSuccess )
			return e;
		fnKernel<<<nCtas, EVAL_THREADS, iDyn, tStream>>> ( P );
		return cudaGetLastError();
	}
	if ( iMode==1 )
		return LaunchStreamT ( stream_kernel<512,1>, P, 1, iDyn, nCtas, tStream );
	if ( iMode==2 )
		return LaunchStreamT ( stream_kernel<512,2>, P, 1, iDyn, nCtas, tStream );
	if ( StreamMiniWidth ( nStack )==512 )
		return LaunchStreamT ( stream_kernel<512,0>, P, nStack, iDyn, nCtas, tStream );
	return LaunchStreamT ( stream_kernel<256,0>, P, nStack, iDyn, nCtas, tStream );
}

template<typename KERNEL>
static int StreamOccupancyT ( KERNEL fnKernel, size_t iDyn )
{
	int n = 0;
	cudaFuncSetAttribute ( fnKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iDyn );
	if ( cudaOccupancyMaxActiveBlocksPerMultiprocessor ( &n, fnKernel, EVAL_THREADS, iDyn )!=cudaSuccess )
		return 1;
	return n>0 ? n : 1;
}

int StreamOccupancy ( int nStack, int iMode )
{
	const size_t iDyn = StreamDynSmemBytes ( nStack, iMode );
	if ( iMode==3 )
		return std::min ( StreamOccupancyT ( orbits_kernel<2>, iDyn ), StreamOccupancyT ( orbits_kernel<4>, iDyn ) );
	if ( iMode==1 )
		return StreamOccupancyT ( stream_kernel<512,1>, iDyn );
	if ( iMode==2 )
		return StreamOccupancyT ( stream_kernel<512,2>, iDyn );
	if ( StreamMiniWidth ( nStack )==512 )
		return StreamOccupancyT ( stream_kernel<512,0>, iDyn );
	return StreamOccupancyT ( stream_kernel<256,0>, iDyn );
}

cudaError_t LaunchAnd ( const EvalParams_t & P, bool bHits, int nCtas, cudaStream_t tStream )
{
	if ( bHits )
		and_kernel<true><<<nCtas, EVAL_THREADS, 0, tStream>>> ( P );
	else
		and_kernel<false><<<nCtas, EVAL_THREADS, 0, tStream>>> ( P );
	return cudaGetLastError();
}

int AndOccupancy ( bool bHits )
{
	int n = 0;
	cudaError_t e = bHits
		? cudaOccupancyMaxActiveBlocksPerMultiprocessor ( &n, and_kernel<true>, EVAL_THREADS, 0 )
		: cudaOccupancyMaxActiveBlocksPerMultiprocessor ( &n, and_kernel<false>, EVAL_THREADS, 0 );
	if ( e!=cudaSuccess )
		return 1;
	return n>0 ? n : 1;
}

cudaError_t LaunchHotDecode ( const HotDecodeParams_t & P, int nCtas, cudaStream_t tStream )
{
	hot_decode_kernel<<<nCtas, EVAL_THREADS, 0, tStream>>> ( P );
	return cudaGetLastError();
}

cudaError_t LaunchSparseDecode ( const SparseDecodeParams_t & P, int nCtas, cudaStream_t tStream )
{
	sparse_decode_kernel<<<nCtas, EVAL_THREADS, 0, tStream>>> ( P );
	return cudaGetLastError();
}

cudaError_t LaunchMerge ( const MergeParams_t & P, int nCtas, cudaStream_t tStream )
{
	merge_kernel<<<nCtas, 256, 0, tStream>>> ( P );
	return cudaGetLastError();
}

cudaError_t LaunchShardMerge ( const Key128_t * pKeys, const int32_t * pCounts, int nShards, int nQueries, int iK,
	Key128_t * pScratch, int iScratchStride, Key128_t * pOutKeys, int32_t * pOutCounts, int nCtas, cudaStream_t tStream,
	const int64_t * pDocids, const uint32_t * pShardBase, int64_t * pOutDocid )
{
	shard_merge_kernel<<<nCtas, 256, 0, tStream>>> ( pKeys, pCounts, nShards, nQueries, iK, pScratch, iScratchStride, pOutKeys, pOutCounts,
		pDocids, pShardBase, pOutDocid );
	return cudaGetLastError();
}

cudaError_t LaunchDecodeDoclist ( const DevIndex_t & tIdx, const DevLeaf_t & tLeaf, uint32_t * pRowid, uint32_t * pHits, uint32_t * pFields,
	uint64_t * pHitlistPos, unsigned long long * pChecksum, int nCtas, cudaStream_t tStream )
{
	decode_doclist_kernel<<<nCtas, EVAL_THREADS, 0, tStream>>> ( tIdx, tLeaf, pRowid, pHits, pFields, pHitlistPos, pChecksum );
	return cudaGetLastError();
}

} // namespace mgpu
