// Hit stage of the fused evaluation kernel (K4/K5/K6): per candidate document, one thread
//   - streams each emitting keyword's hitlist straight out of .spp (DiskIndexQword_c::GetNextHit,
//     src/sphinx.cpp:374-388, 479-501; the single inlined hit of the doclist record, :523-530),
//   - merges the streams in the order the reference's CollectHits() chain produces
//     (ExtAnd_c/ExtOr_c: IsHitLess, src/searchnode.cpp:2611-2615; ExtMultiAnd_T: MergeHits2/3/N :3098-3181;
//     inside a phrase/proximity node: CmpAndHitReverse_fn :2618-2624),
//   - runs the phrase / proximity acceptors (FSMphrase_c::HitFSM :3901-3947, FSMproximity_c::HitFSM :3973-4065),
//   - feeds the ranker state (RankerState_Proximity_fn::Update/Finalize, src/sphinxsearch.cpp:1351-1437;
//     RankerState_Wordcount_fn :1620-1643).
// Everything is a streaming generator: no per-document hit buffer exists, so documents with thousands
// of hits (test_114, test_116) need no special case.  State lives in per-thread local memory.
#pragma once

#include "device_types.h"
#include <stdint.h>

namespace mgpu
{

static const uint32_t EMIT_NWAY_SHIFT = 16;		///< emitter mask: bits 0..15 plain leaves, 16..19 n-way nodes

/// DiskIndexQword_c hit decoder state
struct HitCursor_t
{
	const uint8_t *	m_p;
	const uint8_t *	m_pStart;	///< where the document's hitlist starts in .spp (null: inlined hit): m_p - m_pStart = bytes consumed
	uint32_t		m_uCur;		///< running hitpos (state 0) or the inlined hit (state 1)
	int				m_iState;	///< 0 = stream from .spp, 1 = inlined hit pending, 2 = done
};

__device__ __forceinline__ void SeekHitlist ( HitCursor_t & c, const uint8_t * pSpp, uint64_t uHitlistPos )
{
	if ( uHitlistPos>>63 )
	{
		c.m_iState = 1;
		c.m_uCur = (uint32_t)uHitlistPos;
		c.m_p = nullptr;
		c.m_pStart = nullptr;
	} else
	{
		c.m_iState = 0;
		c.m_uCur = 0;
		c.m_p = pSpp + uHitlistPos;
		c.m_pStart = c.m_p;
	}
}

/// TermAcceptor_T<..>::IsAcceptableHit (src/searchnode.cpp:2264-2285): ^keyword, keyword$, both, @field[N]
__device__ __forceinline__ bool AcceptHit ( uint32_t h, int iTermPos )
{
	const uint32_t uPos = h & 0x7FFFFFu;
	const bool bEnd = ( h>>23 ) & 1u;
	switch ( iTermPos & 7 )
	{
	case 1:		return uPos==1;
	case 2:		return bEnd;
	case 3:		return uPos==1 && bEnd;
	case 4:		return (int)uPos<=( iTermPos>>3 );
	default:	return true;
	}
}

/// next hit that lies in a queried field and passes the keyword's position filter (ExtTerm_T::CollectHits,
/// src/searchnode.cpp:1971-2010; ExtConditional_T :2331-2400); 0 = EMPTY_HIT
__device__ __forceinline__ uint32_t NextHit ( HitCursor_t & c, uint32_t uQueriedFields, int iTermPos=0 )
{
	while ( true )
	{
		uint32_t h;
		if ( c.m_iState==0 )
		{
			uint32_t d = 0, b;
			do { b = __ldg ( c.m_p++ ); d = ( d<<7 ) + ( b & 0x7f ); } while ( b & 0x80 );
			if ( !d )
			{
				c.m_iState = 2;
				return 0;
			}
			c.m_uCur += d;
			h = c.m_uCur;
		} else if ( c.m_iState==1 )
		{
			c.m_iState = 2;
			h = c.m_uCur;
		} else
			return 0;
		const uint32_t f = h>>24;
		if ( f<32 && ( ( uQueriedFields>>f ) & 1u ) && AcceptHit ( h, iTermPos ) )
			return h;
	}
}

/// ExtConditional_T::GetDocsChunk: does the document (hitlist at uHitlistPos) hold an acceptable hit of this keyword?
__device__ __forceinline__ bool HasAcceptableHit ( const uint8_t * pSpp, uint64_t uHitlistPos, uint32_t uQueriedFields, int iTermPos )
{
	HitCursor_t c;
	SeekHitlist ( c, pSpp, uHitlistPos );
	return NextHit ( c, uQueriedFields, iTermPos )!=0;
}

/// one hit as the ranker sees it (ExtHit_t, src/sphinxint.h:725-736)
struct RankHit_t
{
	uint32_t	m_uHitpos;
	uint32_t	m_uQpos;
	uint32_t	m_uSpanlen;
	uint32_t	m_uWeight;
};

/// acceptor + generator state of one hit-level node (the kinds' private parts share their bytes: per-thread local memory)
struct NWayState_t
{
	// FSMphrase_c: m_dStates; FSMproximity_c: m_dProx; ExtOrder_c (BEFORE): m_dVal[0..15] = the longest prefix found so far,
	// m_dVal[16..31] = the most recent attempt (raw hits)
	uint32_t	m_dVal[NWAY_MAX_SPAN+2];	///< phrase: expected hitpos-with-field per state; proximity: last position per query word
	RankHit_t	m_tHead;					///< next folded hit (m_uHitpos==0: exhausted)
	uint32_t	m_uFirstRawHit;				///< raw hit that completed the first match (doc field mask, src/searchnode.cpp:3827-3833)
	bool		m_bAny;
	union
	{
		struct
		{
			uint8_t		m_dTag[NWAY_MAX_SPAN+2];	///< phrase: m_iTagQword per state
			int			m_nStates;
			uint32_t	m_uExpPos, m_uWords;		///< proximity
			int			m_iMinQindex;
		};
		struct
		{
			// FSMmultinear_c, two children (NEAR)
			uint32_t	m_uLastP, m_uFirstHit, m_uChainWeight;
			uint32_t	m_uFirstQpos, m_uFirstNpos;
		};
		struct
		{
			// ExtOrder_c (BEFORE)
			int			m_nLongest, m_nRecent, m_iPosLongest, m_iPosRecent, m_iField;
			uint32_t	m_dEmit[MAX_PHRASE_WORDS];	///< a complete sequence, handed out hit by hit
			int			m_nEmit, m_iEmitNext;
		};
	};
};

__device__ __forceinline__ void ResetFSM ( const DevNWay_t & n, NWayState_t & s )
{
	if ( n.m_eKind<=NWAY_PROXIMITY )
		s.m_nStates = 0;
	if ( n.m_eKind==NWAY_NEAR )
	{
		// FSMmultinear_c::ResetFSM (m_uFirstQpos / m_uFirstNpos are always set by the chain's first hit before they are read)
		s.m_uLastP = 0;
		s.m_uFirstQpos = 65535;
		s.m_uFirstNpos = 0;
		s.m_uFirstHit = 0;
		s.m_uChainWeight = 0;
	} else if ( n.m_eKind==NWAY_BEFORE )
	{
		s.m_nLongest = s.m_nRecent = 0;
		s.m_iPosLongest = s.m_iPosRecent = 0;
		s.m_iField = -1;
		s.m_nEmit = s.m_iEmitNext = 0;
	}
	if ( n.m_eKind==NWAY_PROXIMITY )
	{
		s.m_uExpPos = 0;
		s.m_uWords = 0;
		s.m_iMinQindex = -1;
		for ( int i=0; i<=n.m_iQLen; ++i )
			s.m_dVal[i] = 0xFFFFFFFFu;
	}
}

/// FSMphrase_c::HitFSM, src/searchnode.cpp:3901-3947
__device__ __forceinline__ bool PhraseFSM ( const DevNWay_t & n, NWayState_t & s, uint32_t uHit, int iQpos, RankHit_t & tOut )
{
	const uint32_t uHPF = uHit & ~( 1u<<23 );
	const int iAtom0 = n.m_dAtomPos[0];
	if ( iQpos==iAtom0 && s.m_nStates<NWAY_MAX_SPAN+2 )
	{
		s.m_dTag[s.m_nStates] = 0;
		s.m_dVal[s.m_nStates] = uHPF + (uint32_t)n.m_dQposDelta[0];
		++s.m_nStates;
	}
	for ( int i=s.m_nStates-1; i>=0; --i )
	{
		if ( s.m_dVal[i]<uHPF )
		{
			--s.m_nStates;					// RemoveFast
			s.m_dVal[i] = s.m_dVal[s.m_nStates];
			s.m_dTag[i] = s.m_dTag[s.m_nStates];
			continue;
		}
		int iTag = s.m_dTag[i];
		if ( s.m_dVal[i]==uHPF && n.m_dAtomPos[iTag+1]==iQpos )
		{
			++iTag;
			s.m_dTag[i] = (uint8_t)iTag;
			s.m_dVal[i] = uHPF + (uint32_t)n.m_dQposDelta[iQpos-iAtom0];
		}
		if ( iTag==n.m_nWords-1 )
		{
			const uint32_t uSpan = (uint32_t)n.m_iQLen;
			tOut.m_uHitpos = uHPF-uSpan;
			tOut.m_uQpos = (uint32_t)iAtom0;
			tOut.m_uSpanlen = uSpan+1;
			tOut.m_uWeight = (uint32_t)n.m_nWords;
			s.m_nStates = 0;				// ResetFSM
			return true;
		}
	}
	return false;
}

/// FSMproximity_c::HitFSM, src/searchnode.cpp:3973-4065
__device__ __forceinline__ bool ProximityFSM ( const DevNWay_t & n, NWayState_t & s, uint32_t uHit, int iQpos, RankHit_t & tOut )
{
	const int iQindex = iQpos-n.m_dAtomPos[0];
	const int nProx = n.m_iQLen+1;
	const uint32_t uQLen = (uint32_t)n.m_iQLen;
	uint32_t uHPF = uHit & ~( 1u<<23 );
	if ( s.m_dVal[iQindex]==0xFFFFFFFFu )
		s.m_uWords++;
	s.m_dVal[iQindex] = uHPF;
	if ( uHPF>=s.m_uExpPos || iQindex==s.m_iMinQindex )
	{
		s.m_iMinQindex = iQindex;
		const int iMinPos = (int)( uHPF-uQLen-(uint32_t)n.m_iOpArg );
		for ( int i=0; i<nProx; ++i )
			if ( s.m_dVal[i]!=0xFFFFFFFFu )
			{
				if ( (int)s.m_dVal[i]<=iMinPos )
				{
					s.m_dVal[i] = 0xFFFFFFFFu;
					s.m_uWords--;
					continue;
				}
				if ( s.m_dVal[i]<uHPF )
				{
					s.m_iMinQindex = i;
					uHPF = s.m_dVal[i];
				}
			}
		s.m_uExpPos = s.m_dVal[s.m_iMinQindex] + uQLen + (uint32_t)n.m_iOpArg;
	}
	if ( s.m_uWords!=(uint32_t)n.m_nWords )
		return false;

	// weight = sum over runs of equal (pos - qindex) of (1 + run length - 1), min 1: sort the deltas (insertion sort, <=32 entries)
	int dDeltas[NWAY_MAX_SPAN+1];
	int nDeltas = 0;
	uint32_t uMax = 0;
	for ( int i=0; i<nProx; ++i )
		if ( s.m_dVal[i]!=0xFFFFFFFFu )
		{
			int v = (int)( s.m_dVal[i]-(uint32_t)i );
			int j = nDeltas++;
			while ( j>0 && dDeltas[j-1]>v )
			{
				dDeltas[j] = dDeltas[j-1];
				--j;
			}
			dDeltas[j] = v;
			uMax = max ( uMax, s.m_dVal[i] );
		}
	uint32_t uCurWeight = 0, uWeight = 0;
	int iLast = -2147483647;
	for ( int i=0; i<nDeltas; ++i )
	{
		if ( dDeltas[i]==iLast )
			uCurWeight++;
		else
		{
			uWeight += uCurWeight ? ( 1+uCurWeight ) : 0;
			uCurWeight = 0;
		}
		iLast = dDeltas[i];
	}
	uWeight += uCurWeight ? ( 1+uCurWeight ) : 0;
	if ( !uWeight )
		uWeight = 1;

	const uint32_t uMinPos = s.m_dVal[s.m_iMinQindex];
	tOut.m_uHitpos = uMinPos;
	tOut.m_uQpos = (uint32_t)n.m_dAtomPos[0];
	tOut.m_uSpanlen = uMax-uMinPos+1;
	tOut.m_uWeight = uWeight;

	s.m_dVal[s.m_iMinQindex] = 0xFFFFFFFFu;
	s.m_iMinQindex = -1;
	s.m_uWords--;
	s.m_uExpPos = 0;
	return true;
}

/// FSMmultinear_c::HitFSM, src/searchnode.cpp:4098-4290, for TWO children that are plain keywords (`a NEAR/n b`): every raw hit
/// has match length, span length and weight 1, so the "longer hit at the same position" step back (:4119-4131) and the overlap
/// restart (:4160-4172) cannot fire. The n-way form (three and more children) keeps m_uFirstQpos across documents in the
/// reference (ResetFSM leaves it alone), i.e. its hits depend on the documents seen before: that one stays on the CPU.
/// uNpos = the child's index in the query, uQpos = its query position.
__device__ __forceinline__ bool NearFSM ( const DevNWay_t & n, NWayState_t & s, uint32_t uHit, uint32_t uQpos, uint32_t uNpos, RankHit_t & tOut )
{
	const uint32_t uPos = uHit & ~( 1u<<23 );
	// a second hit at the position of the last one (the same keyword on both sides), :4103-4134
	if ( s.m_uLastP==uPos )
	{
		if ( uNpos<s.m_uFirstNpos )
		{
			s.m_uFirstQpos = uQpos;	// keep the leftmost child of the query
			s.m_uFirstNpos = uNpos;
		}
		return false;
	}
	// too far from the previous hit (or no previous hit): a new chain starts here, :4137-4154
	if ( s.m_uLastP==0 || ( s.m_uLastP + 1 + (uint32_t)n.m_iOpArg )<=uPos )
	{
		s.m_uFirstHit = s.m_uLastP = uPos;
		s.m_uChainWeight = 1;
		s.m_uFirstQpos = uQpos;
		s.m_uFirstNpos = uNpos;
		return false;
	}
	// the same child again: it becomes the head of the chain, :4173-4190
	if ( uNpos==s.m_uFirstNpos )
	{
		if ( s.m_uLastP<uPos )
		{
			s.m_uFirstHit = s.m_uLastP = uPos;
			s.m_uChainWeight = 1;
			s.m_uFirstQpos = uQpos;
			s.m_uFirstNpos = uNpos;
		}
		return false;
	}
	// the other child within reach: emit; the chain shifts to this hit instead of starting over, :4254-4275
	tOut.m_uHitpos = s.m_uFirstHit;
	tOut.m_uWeight = s.m_uChainWeight+1;
	tOut.m_uQpos = min ( s.m_uFirstQpos, uQpos );
	tOut.m_uSpanlen = 2;
	s.m_uFirstHit = s.m_uLastP = uPos;
	s.m_uChainWeight = 1;
	s.m_uFirstQpos = uQpos;
	return true;
}

/// ExtOrder_c::GetMatchingHits, src/searchnode.cpp:4734-4829, one raw hit of child iChild (keywords: span length 1). A complete
/// sequence lands in s.m_dEmit; returns whether one was completed by this hit.
__device__ __forceinline__ bool OrderFSM ( const DevNWay_t & n, NWayState_t & s, uint32_t uHit, int iChild )
{
	const int iHitField = (int)( uHit>>24 ), iHitPos = (int)( uHit & 0x7FFFFFu );
	uint32_t * pLongest = s.m_dVal, * pRecent = s.m_dVal+MAX_PHRASE_WORDS;
	if ( iHitField!=s.m_iField )
	{
		// another field: both trackers start over; only child 0 can seed (and only then the field is remembered)
		s.m_nLongest = s.m_nRecent = 0;
		if ( iChild==0 )
		{
			pLongest[s.m_nLongest++] = uHit;
			s.m_iPosLongest = iHitPos+1;
			s.m_iField = iHitField;
		}
	} else if ( iChild==s.m_nLongest && iHitPos>=s.m_iPosLongest )
	{
		pLongest[s.m_nLongest++] = uHit;
		s.m_iPosLongest = iHitPos+1;
		if ( s.m_nLongest==n.m_nWords )
		{
			for ( int i=0; i<n.m_nWords; ++i )
				s.m_dEmit[i] = pLongest[i];
			s.m_nEmit = n.m_nWords;
			s.m_iEmitNext = 0;
			s.m_nLongest = s.m_nRecent = 0;
			s.m_iPosRecent = s.m_iPosLongest;
			return true;
		}
	} else if ( iChild==0 )
	{
		pRecent[0] = uHit;
		s.m_nRecent = 1;
		s.m_iPosRecent = iHitPos+1;
		if ( !s.m_nLongest )
		{
			pLongest[s.m_nLongest++] = uHit;
			s.m_iPosLongest = iHitPos+1;
		}
	} else if ( iChild==s.m_nRecent && iHitPos>=s.m_iPosRecent )
	{
		pRecent[s.m_nRecent++] = uHit;
		s.m_iPosRecent = iHitPos+1;
		if ( s.m_nRecent==s.m_nLongest )
		{
			for ( int i=0; i<s.m_nRecent; ++i )
				pLongest[i] = pRecent[i];
			s.m_nRecent = 0;
			s.m_iPosLongest = s.m_iPosRecent;
		}
	}
	return false;
}

/// per-document hit machinery: cursors of every keyword + the n-way generators
struct DocHits_t
{
	HitCursor_t	m_dCur[MAX_LEAVES];
	uint32_t	m_dHead[MAX_LEAVES];	///< current filtered hit of each open cursor (0 = exhausted)
	NWayState_t	m_dNWay[MAX_NWAY];
};

static const uint64_t HITPOS_ABSENT = ~0ull;	///< hit scratch: the keyword does not sit on the slot (optional children: NOTNEAR's right side, quorum)

/// advances hit-level node j to its next hit. Phrase / proximity / NEAR (ExtNWay_T::GetDocsChunk inner loop,
/// src/searchnode.cpp:3805-3848): raw hits of the keywords arrive ordered by (hitpos asc, qpos desc) and go through the acceptor.
/// BEFORE (ExtOrder_c::GetMatchingHits :4706-4829): by position, the lowest child wins a tie; a complete sequence is handed out
/// hit by hit. NOTNEAR (ExtNotNear_c::FilterHits :5352-5380): the MUST keyword's hits that no NOT hit follows within N positions.
/// Quorum (ExtQuorum_c::CollectHits :4543-4565): every hit of the keywords on the document, by (position, qpos).
template<int KIND>
__device__ __noinline__ void NWayAdvanceT ( const DevQuery_t & q, int j, DocHits_t & H )
{
	const DevNWay_t & n = q.m_dNWay[j];
	NWayState_t & s = H.m_dNWay[j];
	auto fnEmitted = [&] ( uint32_t uRawHit )
	{
		if ( !s.m_bAny )
		{
			s.m_bAny = true;
			s.m_uFirstRawHit = uRawHit;
		}
	};
	auto fnNext = [&] ( int l )
	{
		H.m_dHead[l] = NextHit ( H.m_dCur[l], q.m_dLeaves[l].m_uQueriedFields, q.m_dLeaves[l].m_iTermPos );
	};
	if constexpr ( KIND==NWAY_NOTNEAR )
	{
		const int lMust = n.m_dLeaf[0], lNot = n.m_dLeaf[1];
		while ( true )
		{
			const uint32_t h = H.m_dHead[lMust];
			if ( !h )
			{
				s.m_tHead.m_uHitpos = 0;
				return;
			}
			const uint32_t uPosMust = h & ~( 1u<<23 );
			while ( H.m_dHead[lNot] && ( H.m_dHead[lNot] & ~( 1u<<23 ) )<uPosMust )
				fnNext ( lNot );	// NOT hits before the MUST hit do not count
			// (the field sits in the top byte, so the distance can be added to the position as it is; a keyword's match length is 1)
			const bool bKeep = !H.m_dHead[lNot] || uPosMust + (uint32_t)n.m_iOpArg<( H.m_dHead[lNot] & ~( 1u<<23 ) );
			fnNext ( lMust );
			if ( bKeep )
			{
				s.m_tHead.m_uHitpos = h; s.m_tHead.m_uQpos = q.m_dLeaves[lMust].m_uAtomPos; s.m_tHead.m_uSpanlen = 1; s.m_tHead.m_uWeight = 1;
				fnEmitted ( h );
				return;
			}
		}
	}
	if constexpr ( KIND!=NWAY_NOTNEAR )
	while ( true )
	{
		if ( KIND==NWAY_BEFORE && s.m_iEmitNext<s.m_nEmit )
		{
			const int i = s.m_iEmitNext++;
			s.m_tHead.m_uHitpos = s.m_dEmit[i]; s.m_tHead.m_uQpos = (uint32_t)n.m_dAtomPos[i]; s.m_tHead.m_uSpanlen = 1; s.m_tHead.m_uWeight = 1;
			fnEmitted ( s.m_dEmit[i] );
			return;
		}
		int iBest = -1, iBestW = 0;
		uint32_t uBestHit = 0, uBestQpos = 0;
		for ( int w=0; w<n.m_nWords; ++w )
		{
			const int l = n.m_dLeaf[w];
			const uint32_t h = H.m_dHead[l];
			if ( !h )
				continue;
			const uint32_t uQpos = q.m_dLeaves[l].m_uAtomPos;
			bool bLess;
			if ( KIND==NWAY_BEFORE )
				bLess = ( h & ~( 1u<<23 ) )<( uBestHit & ~( 1u<<23 ) );
			else if ( KIND==NWAY_QUORUM )
				bLess = ( h & ~( 1u<<23 ) )<( uBestHit & ~( 1u<<23 ) ) || ( ( h & ~( 1u<<23 ) )==( uBestHit & ~( 1u<<23 ) ) && uQpos<uBestQpos );
			else
				bLess = h<uBestHit || ( h==uBestHit && uQpos>uBestQpos );
			if ( iBest<0 || bLess )
			{
				iBest = l; iBestW = w; uBestHit = h; uBestQpos = uQpos;
			}
		}
		if ( iBest<0 )
		{
			s.m_tHead.m_uHitpos = 0;
			return;
		}
		fnNext ( iBest );
		bool bEmit;
		switch ( KIND )
		{
		case NWAY_PROXIMITY:	bEmit = ProximityFSM ( n, s, uBestHit, (int)uBestQpos, s.m_tHead ); break;
		case NWAY_NEAR:			bEmit = NearFSM ( n, s, uBestHit, uBestQpos, (uint32_t)iBestW, s.m_tHead ); break;
		case NWAY_BEFORE:
			OrderFSM ( n, s, uBestHit, iBestW );
			continue;	// (a completed sequence is handed out at the top of the loop)
		case NWAY_QUORUM:
			s.m_tHead.m_uHitpos = uBestHit; s.m_tHead.m_uQpos = uBestQpos; s.m_tHead.m_uSpanlen = 1; s.m_tHead.m_uWeight = 1;
			bEmit = true;
			break;
		default:				bEmit = PhraseFSM ( n, s, uBestHit, (int)uBestQpos, s.m_tHead ); break;
		}
		if ( bEmit )
		{
			fnEmitted ( uBestHit );
			return;
		}
	}
}

/// (one instantiation per kind: the per-hit loops of phrase / proximity carry no tests for the other kinds)
__device__ __forceinline__ void NWayAdvance ( const DevQuery_t & q, int j, DocHits_t & H )
{
	switch ( q.m_dNWay[j].m_eKind )
	{
	case NWAY_PHRASE:		NWayAdvanceT<NWAY_PHRASE> ( q, j, H ); break;
	case NWAY_PROXIMITY:	NWayAdvanceT<NWAY_PROXIMITY> ( q, j, H ); break;
	case NWAY_NEAR:			NWayAdvanceT<NWAY_NEAR> ( q, j, H ); break;
	case NWAY_BEFORE:		NWayAdvanceT<NWAY_BEFORE> ( q, j, H ); break;
	case NWAY_NOTNEAR:		NWayAdvanceT<NWAY_NOTNEAR> ( q, j, H ); break;
	default:				NWayAdvanceT<NWAY_QUORUM> ( q, j, H ); break;
	}
}

/// opens the cursors of hit-level node j on the document in slot `s` (hitlist positions at pHitpos[leaf*iStride+s]) and produces its first hit
__device__ void NWayOpen ( const DevIndex_t & tIdx, const DevQuery_t & q, int j, const uint64_t * pHitpos, int iStride, int s, DocHits_t & H )
{
	const DevNWay_t & n = q.m_dNWay[j];
	for ( int w=0; w<n.m_nWords; ++w )
	{
		const int l = n.m_dLeaf[w];
		const uint64_t uHitpos = pHitpos[(size_t)l*iStride+s];
		if ( uHitpos==HITPOS_ABSENT && n.m_eKind>=NWAY_NOTNEAR )
		{
			H.m_dCur[l].m_iState = 2;
			H.m_dCur[l].m_p = H.m_dCur[l].m_pStart = nullptr;
			H.m_dHead[l] = 0;
			continue;
		}
		SeekHitlist ( H.m_dCur[l], tIdx.m_pSpp, uHitpos );
		H.m_dHead[l] = NextHit ( H.m_dCur[l], q.m_dLeaves[l].m_uQueriedFields, q.m_dLeaves[l].m_iTermPos );
	}
	H.m_dNWay[j].m_bAny = false;
	H.m_dNWay[j].m_uFirstRawHit = 0;
	ResetFSM ( n, H.m_dNWay[j] );
	NWayAdvance ( q, j, H );
}

/// RankerState_Proximity_fn<true,HANDLE_DUPES> + RankerState_Wordcount_fn
struct RankState_t
{
	uint8_t		m_uLCS[MAX_FIELDS];
	uint8_t		m_uCurLCS;
	int			m_iExpDelta;
	int			m_iLastHitPosWithField;
	uint32_t	m_uLcsTailPos, m_uLcsTailQposMask, m_uCurQposMask, m_uCurPos;
	uint32_t	m_uWordcount;				///< WORDCOUNT sum / FIELDMASK bits
	uint32_t	m_uHeadHit, m_uExactHit, m_uMinExpPos;	///< SPH04 (RankerState_ProximityBM25Exact_fn)
	uint8_t		m_uMatchMask[MAX_FIELDS];	///< MATCHANY
};

__device__ __forceinline__ void RankInit ( RankState_t & r, int eRanker )
{
	#pragma unroll
	for ( int i=0; i<MAX_FIELDS; ++i )
		r.m_uLCS[i] = 0;
	r.m_uCurLCS = 0;
	r.m_iExpDelta = eRanker==7 ? -2147483647 : -1;
	r.m_iLastHitPosWithField = -1;
	r.m_uLcsTailPos = 0; r.m_uLcsTailQposMask = 0; r.m_uCurQposMask = 0; r.m_uCurPos = 0;
	r.m_uWordcount = 0;
	r.m_uHeadHit = 0; r.m_uExactHit = 0; r.m_uMinExpPos = 0;
	#pragma unroll
	for ( int i=0; i<MAX_FIELDS; ++i )
		r.m_uMatchMask[i] = 0;
}

__device__ __forceinline__ void RankUpdate ( const DevQuery_t & q, RankState_t & r, const RankHit_t & h )
{
	const uint32_t uField = h.m_uHitpos>>24;	// <32: hits of other fields never get here
	if ( q.m_eRanker==3 )
	{
		r.m_uWordcount += (uint32_t)q.m_dWeights[uField];	// src/sphinxsearch.cpp:1620-1643
		return;
	}
	if ( q.m_eRanker==6 )
	{
		r.m_uWordcount |= 1u<<uField;	// RankerState_Fieldmask_fn, src/sphinxsearch.cpp:1648-1668
		return;
	}
	if ( q.m_eRanker==7 )
	{
		// RankerState_ProximityBM25Exact_fn::Update, src/sphinxsearch.cpp:1476-1514. The state starts every document with
		// m_iExpDelta = -INT_MAX: the reference's stale m_uMinExpPos makes a document's first hit take the same branch.
		const int iPosWithField = (int)( h.m_uHitpos & ~( 1u<<23 ) );
		const int iDelta = iPosWithField - (int)h.m_uQpos;
		const uint32_t uPos = h.m_uHitpos & 0x7FFFFFu;
		const bool bEnd = ( h.m_uHitpos>>23 ) & 1u;
		if ( iDelta==r.m_iExpDelta && (uint32_t)iPosWithField>=r.m_uMinExpPos )
		{
			if ( iPosWithField>r.m_iLastHitPosWithField )
				r.m_uCurLCS = (uint8_t)( r.m_uCurLCS + h.m_uWeight );
			if ( bEnd && (int)h.m_uQpos==q.m_iMaxQpos && (int)uPos==q.m_iMaxQpos )
				r.m_uExactHit |= 1u<<uField;
		} else
		{
			if ( iPosWithField>r.m_iLastHitPosWithField )
				r.m_uCurLCS = (uint8_t)h.m_uWeight;
			if ( uPos==1 )
			{
				r.m_uHeadHit |= 1u<<uField;
				if ( bEnd && q.m_iMaxQpos==1 )
					r.m_uExactHit |= 1u<<uField;
			}
		}
		if ( r.m_uCurLCS>r.m_uLCS[uField] )
			r.m_uLCS[uField] = r.m_uCurLCS;
		r.m_iExpDelta = iDelta + (int)h.m_uSpanlen - 1;
		r.m_iLastHitPosWithField = iPosWithField;
		r.m_uMinExpPos = (uint32_t)iPosWithField + 1;
		return;
	}
	if ( q.m_eRanker==5 )
		r.m_uMatchMask[uField] |= (uint8_t)( 1u<<( ( h.m_uQpos-1 ) & 31u ) );	// RankerState_MatchAny_fn::Update (BYTE mask)
	if ( !q.m_bDupes || q.m_eRanker==5 )
	{
		// src/sphinxsearch.cpp:1357-1367
		const int iPosWithField = (int)( h.m_uHitpos & ~( 1u<<23 ) );
		const int iDelta = iPosWithField - (int)h.m_uQpos;
		if ( iPosWithField>r.m_iLastHitPosWithField )
			r.m_uCurLCS = (uint8_t)( ( ( iDelta==r.m_iExpDelta ) ? r.m_uCurLCS : 0 ) + (uint8_t)h.m_uWeight );
		if ( r.m_uCurLCS>r.m_uLCS[uField] )
			r.m_uLCS[uField] = r.m_uCurLCS;
		r.m_iLastHitPosWithField = iPosWithField;
		r.m_iExpDelta = iDelta + (int)h.m_uSpanlen - 1;
	} else
	{
		// src/sphinxsearch.cpp:1370-1411
		const uint32_t uPos = h.m_uHitpos & ~( 1u<<23 );
		if ( ( r.m_uCurPos>>24 )!=uField )
			r.m_uCurQposMask = 0;
		if ( uPos!=r.m_uCurPos )
		{
			if ( r.m_uCurLCS<2 )
			{
				r.m_uLcsTailPos = r.m_uCurPos;
				r.m_uLcsTailQposMask = r.m_uCurQposMask;
				r.m_uCurLCS = 1;
			}
			r.m_uCurQposMask = 0;
			r.m_uCurPos = uPos;
			if ( r.m_uLCS[uField]<h.m_uWeight )
				r.m_uLCS[uField] = (uint8_t)h.m_uWeight;
		}
		const uint32_t uQposBit = h.m_uQpos<32 ? ( 1u<<h.m_uQpos ) : 0u;	// (DWORD)(1UL<<qpos)
		r.m_uCurQposMask |= uQposBit;
		const int iDelta = (int)( r.m_uCurPos-r.m_uLcsTailPos );
		if ( iDelta>0 && iDelta<32 && ( ( r.m_uCurQposMask>>iDelta ) & r.m_uLcsTailQposMask ) )
		{
			r.m_uLcsTailQposMask = uQposBit;
			r.m_uLcsTailPos = r.m_uCurPos;
			r.m_uCurLCS = (uint8_t)( r.m_uCurLCS+h.m_uWeight );
			r.m_uCurQposMask = 0;
			if ( r.m_uCurLCS>r.m_uLCS[uField] )
				r.m_uLCS[uField] = r.m_uCurLCS;
		}
	}
}

/// Streams the document's hits (root CollectHits order) through the ranker state.
/// uEmit = emitters sitting on this document (plain leaves + n-way nodes). Returns false if the document yields no
/// hits (ExtRanker_State_T skips it, src/sphinxsearch.cpp:1299-1304); else iWeight = final weight before index weight.
/// uHitBytes += the .spp bytes of the document's hitlists that were read (the hitlist share of the algorithmic bytes).
__device__ bool RankDocByHits ( const DevIndex_t & tIdx, const DevQuery_t & q, uint32_t uEmit, const uint64_t * pHitpos, int iStride, int s,
	int iSeedWeight, DocHits_t & H, int & iWeight, unsigned long long & uHitBytes )
{
	uint32_t uLeaves = uEmit & 0xFFFFu;
	uint32_t uNWays = uEmit>>EMIT_NWAY_SHIFT;
	for ( uint32_t m=uLeaves; m; m&=m-1 )
	{
		const int l = __ffs ( m )-1;
		SeekHitlist ( H.m_dCur[l], tIdx.m_pSpp, pHitpos[(size_t)l*iStride+s] );
		H.m_dHead[l] = NextHit ( H.m_dCur[l], q.m_dLeaves[l].m_uQueriedFields, q.m_dLeaves[l].m_iTermPos );
	}
	for ( uint32_t m=uNWays; m; m&=m-1 )
		NWayOpen ( tIdx, q, __ffs ( m )-1, pHitpos, iStride, s, H );

	RankState_t R;
	RankInit ( R, q.m_eRanker );
	bool bAny = false;
	while ( true )
	{
		// k-way merge by (hitpos, qpos); full ties go to the emitter met first in tree order
		int iBest = -1;
		uint32_t uBestHit = 0, uBestQpos = 0;
		for ( uint32_t m=uLeaves; m; m&=m-1 )
		{
			const int l = __ffs ( m )-1;
			const uint32_t h = H.m_dHead[l];
			if ( !h )
				continue;
			const uint32_t uQpos = q.m_dLeaves[l].m_uAtomPos;
			if ( iBest<0 || h<uBestHit || ( h==uBestHit && uQpos<uBestQpos ) )
			{
				iBest = l; uBestHit = h; uBestQpos = uQpos;
			}
		}
		for ( uint32_t m=uNWays; m; m&=m-1 )
		{
			const int j = __ffs ( m )-1;
			const RankHit_t & t = H.m_dNWay[j].m_tHead;
			if ( !t.m_uHitpos )
				continue;
			if ( iBest<0 || t.m_uHitpos<uBestHit || ( t.m_uHitpos==uBestHit && t.m_uQpos<uBestQpos ) )
			{
				iBest = MAX_LEAVES+j; uBestHit = t.m_uHitpos; uBestQpos = t.m_uQpos;
			}
		}
		if ( iBest<0 )
			break;
		bAny = true;
		RankHit_t t;
		if ( iBest<MAX_LEAVES )
		{
			t.m_uHitpos = uBestHit; t.m_uQpos = uBestQpos; t.m_uSpanlen = 1; t.m_uWeight = 1;
			H.m_dHead[iBest] = NextHit ( H.m_dCur[iBest], q.m_dLeaves[iBest].m_uQueriedFields, q.m_dLeaves[iBest].m_iTermPos );
		} else
		{
			t = H.m_dNWay[iBest-MAX_LEAVES].m_tHead;
			NWayAdvance ( q, iBest-MAX_LEAVES, H );
		}
		RankUpdate ( q, R, t );
	}
	if ( !bAny )
		return false;
	{
		uint32_t uAllLeaves = uLeaves;
		for ( uint32_t m=uNWays; m; m&=m-1 )
		{
			const DevNWay_t & n = q.m_dNWay[__ffs ( m )-1];
			for ( int w=0; w<n.m_nWords; ++w )
				uAllLeaves |= 1u<<n.m_dLeaf[w];
		}
		for ( uint32_t m=uAllLeaves; m; m&=m-1 )
		{
			const HitCursor_t & c = H.m_dCur[__ffs ( m )-1];
			if ( c.m_pStart )
				uHitBytes += (unsigned long long)( c.m_p-c.m_pStart );
		}
	}

	if ( q.m_eRanker==3 || q.m_eRanker==6 )
		iWeight = (int)R.m_uWordcount;
	else if ( q.m_eRanker==7 )
	{
		// RankerState_ProximityBM25Exact_fn::Finalize, src/sphinxsearch.cpp:1516-1535
		uint32_t uRank = 0;
		for ( int i=0; i<q.m_nWeights; ++i )
			uRank += ( 4u*R.m_uLCS[i] + 2u*( ( R.m_uHeadHit>>i ) & 1u ) + ( ( R.m_uExactHit>>i ) & 1u ) )*(uint32_t)q.m_dWeights[i];
		iWeight = (int)( (uint32_t)iSeedWeight + uRank*1000u );
	} else if ( q.m_eRanker==5 )
	{
		// RankerState_MatchAny_fn::Finalize, src/sphinxsearch.cpp:1604-1621
		uint32_t uPhraseK = 0;
		for ( int i=0; i<q.m_nWeights; ++i )
			uPhraseK += (uint32_t)q.m_dWeights[i]*(uint32_t)q.m_nQwords;
		uint32_t uRank = 0;
		for ( int i=0; i<q.m_nWeights; ++i )
			if ( R.m_uMatchMask[i] )
				uRank += ( (uint32_t)__popc ( R.m_uMatchMask[i] ) + ( (uint32_t)R.m_uLCS[i]-1u )*uPhraseK )*(uint32_t)q.m_dWeights[i];
		iWeight = (int)uRank;
	} else if ( q.m_eRanker==4 )
	{
		// RankerState_Proximity_fn<false,..>::Finalize: the bare rank
		uint32_t uRank = 0;
		for ( int i=0; i<q.m_nWeights; ++i )
			uRank += (uint32_t)R.m_uLCS[i]*(uint32_t)q.m_dWeights[i];
		iWeight = (int)uRank;
	} else
	{
		// Finalize, src/sphinxsearch.cpp:1415-1437
		uint32_t uRank = 0;
		for ( int i=0; i<q.m_nWeights; ++i )
			uRank += (uint32_t)R.m_uLCS[i]*(uint32_t)q.m_dWeights[i];
		iWeight = (int)( (uint32_t)iSeedWeight + uRank*1000u );
	}
	return true;
}

} // namespace mgpu
