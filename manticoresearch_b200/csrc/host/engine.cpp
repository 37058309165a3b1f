// Index loader, query planner and batch executor of the GPU query engine. See engine.h.
#include "engine.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mgpu
{

static const int AND_CHUNK = 128;	///< = AND_CHUNK_BLOCKS of kernels.cu: driver blocks per CTA round

#define CUDA_TRY(_expr,_err) \
	do { cudaError_t _e = (_expr); if ( _e!=cudaSuccess ) { _err = std::string ( #_expr ": " ) + cudaGetErrorString ( _e ); return MGPU_E_CUDA; } } while (0)

//////////////////////////////////////////////////////////////////////////
// loader
//////////////////////////////////////////////////////////////////////////

namespace
{

struct MappedFile_t
{
	const BYTE * m_p = nullptr;
	size_t m_iLen = 0;
	bool Map ( const std::string & sPath )
	{
		int fd = open ( sPath.c_str(), O_RDONLY );
		if ( fd<0 )
			return false;
		struct stat st;
		if ( fstat ( fd, &st )<0 ) { close ( fd ); return false; }
		m_iLen = (size_t)st.st_size;
		if ( m_iLen )
		{
			void * p = mmap ( nullptr, m_iLen, PROT_READ, MAP_PRIVATE, fd, 0 );
			if ( p==MAP_FAILED ) { close ( fd ); return false; }
			m_p = (const BYTE *)p;
		}
		close ( fd );
		return true;
	}
	~MappedFile_t() { if ( m_p ) munmap ( (void*)m_p, m_iLen ); }
};

template<typename T>
int Upload ( DevBuf_T<T> & tDst, const void * pSrc, size_t nBytes, size_t nPadBytes, std::string & sError )
{
	size_t nElems = ( nBytes+nPadBytes+sizeof(T)-1 )/sizeof(T);
	CUDA_TRY ( tDst.Alloc ( nElems ), sError );
	CUDA_TRY ( cudaMemset ( tDst.m_p, 0, nElems*sizeof(T) ), sError );
	if ( nBytes )
		CUDA_TRY ( cudaMemcpy ( tDst.m_p, pSrc, nBytes, cudaMemcpyHostToDevice ), sError );
	return MGPU_OK;
}

} // namespace


void * Index_c::Pinned ( size_t nBytes )
{
	if ( nBytes>m_nPinned )
	{
		if ( m_pPinned )
			cudaFreeHost ( m_pPinned );
		m_pPinned = nullptr;
		m_nPinned = 0;
		if ( cudaHostAlloc ( &m_pPinned, nBytes, cudaHostAllocDefault )!=cudaSuccess )
			return nullptr;
		m_nPinned = nBytes;
	}
	return m_pPinned;
}

void * Index_c::PinnedUpload ( size_t nBytes )
{
	if ( nBytes>m_nPinnedUp )
	{
		if ( m_pPinnedUp )
			cudaFreeHost ( m_pPinnedUp );
		m_pPinnedUp = nullptr;
		m_nPinnedUp = 0;
		if ( cudaHostAlloc ( &m_pPinnedUp, nBytes+nBytes/4, cudaHostAllocDefault )!=cudaSuccess )
			return nullptr;
		m_nPinnedUp = nBytes+nBytes/4;
	}
	return m_pPinnedUp;
}

Index_c::~Index_c()
{
	if ( m_pPinnedUp )
		cudaFreeHost ( m_pPinnedUp );
	if ( m_pPinned )
		cudaFreeHost ( m_pPinned );
	if ( m_tOwnStream )
	{
		cudaSetDevice ( m_iDevice );
		cudaStreamDestroy ( m_tOwnStream );
	}
	if ( m_tHotStream )
	{
		cudaSetDevice ( m_iDevice );
		cudaStreamDestroy ( m_tHotStream );
	}
}

int Index_c::AttrIndex ( const char * szName ) const
{
	for ( size_t i=0; i<m_tHdr.m_dAttrs.size(); ++i )
		if ( m_tHdr.m_dAttrs[i].m_sName==szName )
			return (int)i;
	return -1;
}

int Index_c::FieldIndex ( const char * szName ) const
{
	for ( size_t i=0; i<m_tHdr.m_dFields.size(); ++i )
		if ( m_tHdr.m_dFields[i].m_sName==szName )
			return (int)i;
	return -1;
}

int Index_c::Open ( const char * szPrefix, int iDevice, uint32_t uRowidBase )
{
	const std::string sPrefix ( szPrefix );
	m_iDevice = iDevice;
	m_uRowidBase = uRowidBase;

	// the product has no CPU path: fail loudly without a device
	int nDevices = 0;
	if ( cudaGetDeviceCount ( &nDevices )!=cudaSuccess || nDevices<=0 )
	{
		m_sError = "no CUDA device available (the GPU query path has no CPU fallback)";
		return MGPU_E_NO_DEVICE;
	}
	if ( iDevice<0 || iDevice>=nDevices )
	{
		m_sError = "CUDA device ordinal out of range";
		return MGPU_E_NO_DEVICE;
	}
	CUDA_TRY ( cudaSetDevice ( iDevice ), m_sError );
	cudaDeviceProp tProp;
	CUDA_TRY ( cudaGetDeviceProperties ( &tProp, iDevice ), m_sError );
	m_nSMs = tProp.multiProcessorCount;

	MappedFile_t tSph, tSpi, tSpd, tSpp, tSpe, tSpa, tSpm;
	if ( !tSph.Map ( sPrefix+".sph" ) || !tSpi.Map ( sPrefix+".spi" ) || !tSpd.Map ( sPrefix+".spd" )
		|| !tSpp.Map ( sPrefix+".spp" ) || !tSpe.Map ( sPrefix+".spe" ) || !tSpa.Map ( sPrefix+".spa" ) )
	{
		m_sError = "failed to open index files at " + sPrefix + ".{sph,spi,spd,spp,spe,spa}";
		return MGPU_E_IO;
	}
	tSpm.Map ( sPrefix+".spm" );

	if ( !ReadHeader ( tSph.m_p, tSph.m_iLen, m_tHdr, m_sError ) )
		return MGPU_E_FORMAT;
	if ( m_tHdr.m_eHitless!=SPH_HITLESS_NONE )
	{
		m_sError = "hitless indexes are not supported";
		return MGPU_E_FORMAT;
	}
	if ( m_tHdr.m_iSkiplistBlockSize!=32 )
	{
		m_sError = "skiplist_block_size must be 32 (one warp per skiplist block)";
		return MGPU_E_FORMAT;
	}
	if ( m_tHdr.m_dFields.size()>MAX_FIELDS )
	{
		m_sError = "more than 32 full-text fields are not supported";
		return MGPU_E_FORMAT;
	}
	if ( m_tHdr.m_iDocinfo>=(int64_t)0xFFFF0000u )
	{
		m_sError = "too many rows";
		return MGPU_E_FORMAT;
	}
	const int iStride = m_tHdr.RowStride();
	if ( (int64_t)tSpa.m_iLen<m_tHdr.m_iDocinfo*iStride*4 )
	{
		m_sError = ".spa is shorter than the header says";
		return MGPU_E_FORMAT;
	}

	std::vector<DictEntry_t> dDict;
	if ( !ReadDictionary ( tSpi.m_p, tSpi.m_iLen, m_tHdr, dDict, m_sError ) )
		return MGPU_E_FORMAT;

	// doclist extents: doclists are laid back to back in .spd, so the length is the distance to the next one
	{
		std::vector<size_t> dOrder ( dDict.size() );
		for ( size_t i=0; i<dOrder.size(); ++i )
			dOrder[i] = i;
		std::sort ( dOrder.begin(), dOrder.end(), [&] ( size_t a, size_t b ) { return dDict[a].m_iDoclistOffset<dDict[b].m_iDoclistOffset; } );
		for ( size_t k=0; k<dOrder.size(); ++k )
		{
			DictEntry_t & e = dDict[dOrder[k]];
			int64_t iNext = k+1<dOrder.size() ? dDict[dOrder[k+1]].m_iDoclistOffset : (int64_t)tSpd.m_iLen;
			e.m_iDoclistLength = iNext-e.m_iDoclistOffset;
			if ( e.m_iDoclistOffset<=0 || e.m_iDoclistLength<=0 || iNext>(int64_t)tSpd.m_iLen )
			{
				m_sError = "dictionary entry points outside .spd";
				return MGPU_E_FORMAT;
			}
			// a doclist record is four varints, the list ends with a zero byte; a keyword cannot sit in more documents than the index has
			if ( e.m_iDocs<1 || (int64_t)e.m_iDocs>m_tHdr.m_iDocinfo || (int64_t)e.m_iDocs*4+1>e.m_iDoclistLength )
			{
				m_sError = "implausible document count for keyword " + e.m_sKeyword;
				return MGPU_E_FORMAT;
			}
		}
	}

	// decode every skiplist once into the flat block table (DiskIndexQwordSetup_c::Setup, src/sphinx.cpp:13056-13073,
	// except that we keep the trailing partial block's entry, which the writer emits (:8447-8453) and the reader skips)
	std::vector<uint32_t> dBlkRowid;
	std::vector<uint64_t> dBlkOff, dBlkHitpos;
	size_t nTotalBlocks = 0;
	for ( const auto & e : dDict )
		nTotalBlocks += ( (size_t)e.m_iDocs+31 )/32;
	dBlkRowid.reserve ( nTotalBlocks );
	dBlkOff.reserve ( nTotalBlocks );
	dBlkHitpos.reserve ( nTotalBlocks );
	m_hTerms.reserve ( dDict.size()*2 );
	for ( const auto & e : dDict )
	{
		TermInfo_t t;
		t.m_uFirstBlk = (uint32_t)dBlkRowid.size();
		t.m_nBlocks = (uint32_t)( ( e.m_iDocs+31 )/32 );
		t.m_iDocs = e.m_iDocs;
		t.m_iHits = e.m_iHits;
		t.m_iDoclistOffset = e.m_iDoclistOffset;
		t.m_iDoclistLength = e.m_iDoclistLength;

		uint32_t uRow = 0;
		uint64_t uOff = (uint64_t)e.m_iDoclistOffset, uHit = 0;
		dBlkRowid.push_back ( uRow ); dBlkOff.push_back ( uOff ); dBlkHitpos.push_back ( uHit );
		if ( e.m_iDocs>32 )
		{
			if ( e.m_iSkiplistOffset<=0 || e.m_iSkiplistOffset>=(int64_t)tSpe.m_iLen )
			{
				m_sError = "skiplist offset outside .spe";
				return MGPU_E_FORMAT;
			}
			ByteReader_t r ( tSpe.m_p+e.m_iSkiplistOffset, tSpe.m_iLen-e.m_iSkiplistOffset );
			for ( uint32_t b=1; b<t.m_nBlocks; ++b )
			{
				const uint64_t uRowNext = (uint64_t)uRow + 32 + r.Unzip();
				if ( uRowNext>(uint64_t)m_tHdr.m_iDocinfo )
				{
					m_sError = "skiplist rowid beyond the index for keyword " + e.m_sKeyword;
					return MGPU_E_FORMAT;
				}
				uRow = (uint32_t)uRowNext;
				uOff += 4*32 + r.Unzip();
				uHit += r.Unzip();
				dBlkRowid.push_back ( uRow ); dBlkOff.push_back ( uOff ); dBlkHitpos.push_back ( uHit );
			}
			if ( r.m_bError || uOff>=(uint64_t)( e.m_iDoclistOffset+e.m_iDoclistLength ) )
			{
				m_sError = "corrupt skiplist for keyword " + e.m_sKeyword;
				return MGPU_E_FORMAT;
			}
			t.m_iSkiplistBytes = (int64_t)( r.m_p-( tSpe.m_p+e.m_iSkiplistOffset ) );
		}
		t.m_iOrdinal = (int)m_hTerms.size();
		m_hTerms.emplace ( e.m_sKeyword, t );
	}
	m_dTermUse.assign ( m_hTerms.size()+1, 0 );

	// upload
	int iRes;
	if ( ( iRes = Upload ( m_dSpd, tSpd.m_p, tSpd.m_iLen, 64, m_sError ) )!=MGPU_OK ) return iRes;
	if ( ( iRes = Upload ( m_dSpp, tSpp.m_p, tSpp.m_iLen, 64, m_sError ) )!=MGPU_OK ) return iRes;
	if ( ( iRes = Upload ( m_dSpa, tSpa.m_p, (size_t)m_tHdr.m_iDocinfo*iStride*4, 64, m_sError ) )!=MGPU_OK ) return iRes;
	if ( ( iRes = Upload ( m_dBlkRowid, dBlkRowid.data(), dBlkRowid.size()*4, 256, m_sError ) )!=MGPU_OK ) return iRes;
	if ( ( iRes = Upload ( m_dBlkOff, dBlkOff.data(), dBlkOff.size()*8, 64, m_sError ) )!=MGPU_OK ) return iRes;
	if ( ( iRes = Upload ( m_dBlkHitpos, dBlkHitpos.data(), dBlkHitpos.size()*8, 64, m_sError ) )!=MGPU_OK ) return iRes;
	bool bAnyDead = false;
	for ( size_t i=0; i<tSpm.m_iLen && !bAnyDead; ++i )
		bAnyDead = tSpm.m_p[i]!=0;
	if ( bAnyDead )
	{
		std::vector<uint32_t> dDead ( (size_t)( m_tHdr.m_iDocinfo+31 )/32, 0 );
		memcpy ( dDead.data(), tSpm.m_p, std::min ( tSpm.m_iLen, dDead.size()*4 ) );
		if ( ( iRes = Upload ( m_dDead, dDead.data(), dDead.size()*4, 64, m_sError ) )!=MGPU_OK ) return iRes;
	}

	m_tDev.m_pSpd = m_dSpd.m_p;
	m_tDev.m_pSpp = m_dSpp.m_p;
	m_tDev.m_pSpa = m_dSpa.m_p;
	m_tDev.m_pDead = m_dDead.m_p;
	m_tDev.m_pBlkRowid = m_dBlkRowid.m_p;
	m_tDev.m_pBlkOff = m_dBlkOff.m_p;
	m_tDev.m_pBlkHitpos = m_dBlkHitpos.m_p;
	m_tDev.m_iSpdLen = (int64_t)tSpd.m_iLen;
	m_tDev.m_iSppLen = (int64_t)tSpp.m_iLen;
	m_tDev.m_uRows = (uint32_t)m_tHdr.m_iDocinfo;
	m_tDev.m_iStride = iStride;
	m_tDev.m_bInlineHits = ( m_tHdr.m_eHitFormat==SPH_HIT_FORMAT_INLINE ) ? 1 : 0;
	m_tDev.m_uRowidBase = uRowidBase;

	CUDA_TRY ( cudaStreamCreateWithFlags ( &m_tOwnStream, cudaStreamNonBlocking ), m_sError );
	CUDA_TRY ( cudaStreamCreateWithFlags ( &m_tHotStream, cudaStreamNonBlocking ), m_sError );
	m_tStream = m_tOwnStream;
	{
		// per-batch buffers come from the device's stream-ordered pool: keep freed memory cached instead of returning it to the driver
		cudaMemPool_t tPool;
		if ( cudaDeviceGetDefaultMemPool ( &tPool, iDevice )==cudaSuccess )
		{
			uint64_t uKeep = ~0ull;
			cudaMemPoolSetAttribute ( tPool, cudaMemPoolAttrReleaseThreshold, &uKeep );
		}
	}
	return MGPU_OK;
}

//////////////////////////////////////////////////////////////////////////
// planner
//////////////////////////////////////////////////////////////////////////

namespace
{

// sphSort for <=33 elements is this (unstable) insertion sort, src/sphinxstd.h:853-866
template<typename T, typename LESS>
void RefSort ( std::vector<T> & d, LESS fnLess )
{
	for ( size_t i=1; i<d.size(); ++i )
		for ( size_t j=i; j>0; --j )
		{
			if ( fnLess ( d[j-1], d[j] ) )
				break;
			std::swap ( d[j], d[j-1] );
		}
}

enum { PN_TERM, PN_MULTIAND, PN_AND, PN_OR, PN_MAYBE, PN_ANDNOT, PN_NWAY, PN_MULTIOR };

struct PLeaf_t
{
	int					m_iWord = 0;
	const TermInfo_t *	m_pTerm = nullptr;
	uint32_t			m_uFields = 0xFFFFFFFFu;
	int					m_iAtomPos = 0;
	int					m_iNodePos = 0;
	bool				m_bNotWeighted = false;
	bool				m_bOwnsIDF = false;		///< first occurrence of the word in eval-tree order
	int					m_iTermPos = 0;
	int64_t				m_iOrderDocs = 0;		///< what the reference's GetDocsCount() order sees: the caller's global df when given (shards), else the dictionary's
	int64_t				Docs() const { return m_iOrderDocs; }
	int					LocalDocs() const { return m_pTerm ? m_pTerm->m_iDocs : 0; }
};

struct PNode_t
{
	int					m_eKind = PN_TERM;
	int					m_iLeaf = -1;
	std::vector<int>	m_dLeaves;		///< multi-AND: sorted order; n-way: chain order (NOTNEAR: must, not; quorum: children by query position)
	std::vector<int>	m_dRegOrder;	///< hit-level nodes: the order GetQwords() registers the keywords in, when it is not the chain order
	int					m_iLeft = -1, m_iRight = -1;
	int					m_iNWay = -1;
};

struct Planner_c
{
	const Index_c &		m_tIndex;
	const mgpu_query &	m_q;
	PlannedQuery_t &	m_tOut;
	DevQuery_t			m_tFull;			///< the query as built here; Finish() hands the core to m_tOut and the extension to the caller
	std::vector<PLeaf_t> m_dLeaves;
	std::vector<PNode_t> m_dNodes;
	int					m_iError = MGPU_OK;
	int					m_iMaxSp = 0;
	bool				m_bAnyTermPos = false;
	int					m_nVisits = 0;			///< Create() calls: a well-formed tree visits every node once

	struct Qword_t { int m_iDocs; float m_fBoost; int m_iFirstWord; float m_fIDF; };
	std::unordered_map<std::string,Qword_t> m_hQwords;

	Planner_c ( const Index_c & tIndex, const mgpu_query & q, PlannedQuery_t & tOut ) : m_tIndex ( tIndex ), m_q ( q ), m_tOut ( tOut ) {}

	int Fail ( int iErr )	{ if ( m_iError==MGPU_OK ) m_iError = iErr; return -1; }

	int AddLeaf ( const mgpu_xqnode & tNode, int iWord, int iNodePos )
	{
		if ( iWord<0 || iWord>=m_q.n_words || !m_q.words[iWord].word )
			return Fail ( MGPU_E_BAD_QUERY );
		const mgpu_xqkeyword & tWord = m_q.words[iWord];
		if ( (int)m_dLeaves.size()>=MAX_LEAVES )
			return Fail ( MGPU_E_UNSUPPORTED );
		PLeaf_t t;
		t.m_iWord = iWord;
		t.m_pTerm = m_tIndex.FindTerm ( tWord.word );
		// a rowid-range shard must order its keywords (and with them the fp32 TF*IDF additions) like the unsharded index would
		t.m_iOrderDocs = ( m_q.shard_of_global && m_q.word_docs && m_q.word_docs[iWord]>=0 ) ? m_q.word_docs[iWord] : t.LocalDocs();
		t.m_uFields = tNode.field_mask;
		t.m_iAtomPos = tWord.atom_pos;
		t.m_iNodePos = iNodePos;
		t.m_bNotWeighted = tNode.not_weighted!=0;
		// ExtTermPos_T: ^keyword / keyword$ / @field[N] (src/searchnode.cpp:875-878, 1145-1146)
		t.m_iTermPos = ( tWord.field_start && tWord.field_end ) ? 3 : tWord.field_start ? 1 : tWord.field_end ? 2 : 0;
		if ( tNode.field_max_pos )
		{
			if ( tNode.field_max_pos<0 || tNode.field_max_pos>=( 1<<23 ) )
				return Fail ( MGPU_E_BAD_QUERY );
			t.m_iTermPos = 4 | ( tNode.field_max_pos<<3 );
		}
		m_bAnyTermPos |= t.m_iTermPos!=0;
		m_dLeaves.push_back ( t );
		return (int)m_dLeaves.size()-1;
	}

	int NewNode ( const PNode_t & t )	{ m_dNodes.push_back ( t ); return (int)m_dNodes.size()-1; }

	/// ExtNode_i::Create, src/searchnode.cpp:1599-1811
	int Create ( int iNode )
	{
		if ( iNode<0 || iNode>=m_q.n_nodes )
			return Fail ( MGPU_E_BAD_QUERY );
		// a child list naming its own node or an ancestor (or a DAG blown up into an exponential tree) must not recurse forever
		if ( ++m_nVisits>m_q.n_nodes )
			return Fail ( MGPU_E_BAD_QUERY );
		const mgpu_xqnode & tNode = m_q.nodes[iNode];
		if ( tNode.n_words )
		{
			if ( tNode.n_words==1 )
			{
				PNode_t t;
				t.m_eKind = PN_TERM;
				t.m_iLeaf = AddLeaf ( tNode, tNode.first_word, 0 );
				return t.m_iLeaf<0 ? -1 : NewNode ( t );
			}
			if ( tNode.op==MGPU_OP_QUORUM )
			{
				// degenerate quorums (src/searchnode.cpp:1638-1688): threshold >= words -> AND, threshold 1 -> OR, over the keywords sorted by
				// doc count (chains of ExtAnd_c / ExtOr_c); everything else is a real ExtQuorum_c
				const int iCount = tNode.n_words, iThr = tNode.oparg;
				const bool bOr = ( iThr<iCount && iCount<=256 && iThr==1 );
				if ( tNode.first_word<0 || tNode.first_word+iCount>m_q.n_words )
					return Fail ( MGPU_E_BAD_QUERY );
				if ( iThr<iCount && iCount<=256 && iThr!=1 )
					return CreateQuorumNode ( tNode );
				PNode_t t;
				t.m_eKind = bOr ? PN_MULTIOR : PN_MULTIAND;
				for ( int i=0; i<iCount; ++i )
				{
					int iLeaf = AddLeaf ( tNode, tNode.first_word+i, i );
					if ( iLeaf<0 )
						return -1;
					t.m_dLeaves.push_back ( iLeaf );
				}
				RefSort ( t.m_dLeaves, [this] ( int a, int b ) { return m_dLeaves[a].Docs()<m_dLeaves[b].Docs(); } );
				return NewNode ( t );
			}
			if ( tNode.op!=MGPU_OP_PHRASE && tNode.op!=MGPU_OP_PROXIMITY )
				return Fail ( MGPU_E_UNSUPPORTED );
			return CreateMultiNode ( tNode );
		}
		const int nChildren = tNode.n_children;
		if ( nChildren<1 )
			return -1;
		if ( tNode.first_child<0 || tNode.first_child+nChildren>m_q.n_children )
			return Fail ( MGPU_E_BAD_QUERY );
		const int32_t * pChildren = m_q.children + tNode.first_child;
		for ( int i=0; i<nChildren; ++i )
			if ( pChildren[i]<0 || pChildren[i]>=m_q.n_nodes )
				return Fail ( MGPU_E_BAD_QUERY );

		if ( tNode.op==MGPU_OP_NEAR || tNode.op==MGPU_OP_BEFORE || tNode.op==MGPU_OP_NOTNEAR )
			return CreateKeywordOpNode ( tNode, pChildren, nChildren );

		bool bAndTerms = ( tNode.op==MGPU_OP_AND );
		for ( int i=0; i<nChildren && bAndTerms; ++i )
			bAndTerms = ( m_q.nodes[pChildren[i]].n_words==1 );
		if ( bAndTerms && nChildren>1 )
		{
			// ExtMultiAnd_T: children sorted by doc count (src/searchnode.cpp:2791)
			PNode_t t;
			t.m_eKind = PN_MULTIAND;
			for ( int i=0; i<nChildren; ++i )
			{
				const mgpu_xqnode & tChild = m_q.nodes[pChildren[i]];
				int iLeaf = AddLeaf ( tChild, tChild.first_word, i );
				if ( iLeaf<0 )
					return -1;
				t.m_dLeaves.push_back ( iLeaf );
			}
			RefSort ( t.m_dLeaves, [this] ( int a, int b ) { return m_dLeaves[a].Docs()<m_dLeaves[b].Docs(); } );
			return NewNode ( t );
		}
		if ( bAndTerms )
			return Create ( pChildren[0] );

		// the one-child NOT wrapper FixupNots leaves under an ANDNOT: the generic fold of a single child is the child (:1785-1806)
		if ( tNode.op==MGPU_OP_NOT && nChildren==1 )
			return Create ( pChildren[0] );

		int iKind;
		switch ( tNode.op )
		{
		case MGPU_OP_OR:		iKind = PN_OR; break;
		case MGPU_OP_MAYBE:		iKind = PN_MAYBE; break;
		case MGPU_OP_AND:		iKind = PN_AND; break;
		case MGPU_OP_ANDNOT:	iKind = PN_ANDNOT; break;
		default:				return Fail ( MGPU_E_UNSUPPORTED );
		}
		int iCur = -1;
		for ( int i=0; i<nChildren; ++i )
		{
			int iNext = Create ( pChildren[i] );
			if ( m_iError!=MGPU_OK )
				return -1;
			if ( iNext<0 ) continue;
			if ( iCur<0 ) { iCur = iNext; continue; }
			PNode_t t;
			t.m_eKind = iKind;
			t.m_iLeft = iCur;
			t.m_iRight = iNext;
			iCur = NewNode ( t );
		}
		return iCur;
	}

	/// CreateMultiNode<ExtPhrase_c|ExtProximity_c> (plain keywords branch, src/searchnode.cpp:986-1042) + ExtNWay_T::ConstructNode (:3767-3802):
	/// an AND chain over the keywords sorted by doc count, then the acceptor over their merged hits
	int CreateMultiNode ( const mgpu_xqnode & tNode )
	{
		DevQuery_t & d = m_tFull;
		if ( d.m_nNWay>=MAX_NWAY || tNode.n_words>MAX_PHRASE_WORDS )
			return Fail ( MGPU_E_UNSUPPORTED );
		if ( tNode.first_word<0 || tNode.first_word+tNode.n_words>m_q.n_words )
			return Fail ( MGPU_E_BAD_QUERY );
		PNode_t t;
		t.m_eKind = PN_NWAY;
		t.m_iNWay = d.m_nNWay;
		DevNWay_t & n = d.m_dNWay[d.m_nNWay++];
		n.m_eKind = ( tNode.op==MGPU_OP_PROXIMITY ) ? NWAY_PROXIMITY : NWAY_PHRASE;
		n.m_iOpArg = tNode.oparg;
		n.m_nWords = tNode.n_words;
		for ( int i=0; i<tNode.n_words; ++i )
		{
			int iLeaf = AddLeaf ( tNode, tNode.first_word+i, i );
			if ( iLeaf<0 )
				return -1;
			t.m_dLeaves.push_back ( iLeaf );
			n.m_dLeaf[i] = iLeaf;
			n.m_dAtomPos[i] = m_dLeaves[iLeaf].m_iAtomPos;
			if ( i && n.m_dAtomPos[i]<=n.m_dAtomPos[i-1] )
				return Fail ( MGPU_E_BAD_QUERY );
		}
		n.m_dAtomPos[tNode.n_words] = -1;
		n.m_iQLen = n.m_dAtomPos[tNode.n_words-1]-n.m_dAtomPos[0];
		if ( n.m_iQLen>NWAY_MAX_SPAN || n.m_dAtomPos[0]<0 || n.m_dAtomPos[tNode.n_words-1]>0xFFFF )
			return Fail ( MGPU_E_UNSUPPORTED );
		// FSMphrase_c ctor, src/searchnode.cpp:3884-3899
		for ( int i=0; i<=NWAY_MAX_SPAN; ++i )
			n.m_dQposDelta[i] = -INT_MAX;
		for ( int i=1; i<tNode.n_words; ++i )
			n.m_dQposDelta [ n.m_dAtomPos[i-1]-n.m_dAtomPos[0] ] = n.m_dAtomPos[i]-n.m_dAtomPos[i-1];
		// chain order: by doc count (sphSort over positions, :3800)
		RefSort ( t.m_dLeaves, [this] ( int a, int b ) { return m_dLeaves[a].Docs()<m_dLeaves[b].Docs(); } );
		return NewNode ( t );
	}

	/// NEAR / BEFORE / NOTNEAR whose children are plain keywords: hit-level nodes in the n-way slot of the hit stage.
	/// NEAR = CreateMultiNode<ExtMultinear_c> + ExtNWay_T::ConstructNode (src/searchnode.cpp:933-978, 3767-3802): AND chain over the
	/// children sorted by doc count, acceptor FSMmultinear_c; only the two-children form (the n-way FSM carries m_uFirstQpos from one
	/// document to the next in the reference, :4193-4229 vs ResetFSM, so its hits depend on the documents seen before).
	/// BEFORE = CreateOrderNode / ExtOrder_c (:1044-1074, 4657-4935); NOTNEAR = ExtNotNear_c (:5325-5478).
	/// Children that are phrases, OR groups or other operators are not on the GPU path.
	int CreateKeywordOpNode ( const mgpu_xqnode & tNode, const int32_t * pChildren, int nChildren )
	{
		DevQuery_t & d = m_tFull;
		if ( tNode.op!=MGPU_OP_NOTNEAR && nChildren<2 )
			return -1;	// ("order node requires at least two children" / no phrase node from one child: an empty node)
		if ( tNode.op==MGPU_OP_NOTNEAR && nChildren!=2 )
			return Fail ( MGPU_E_BAD_QUERY );
		for ( int i=0; i<nChildren; ++i )
		{
			const mgpu_xqnode & tChild = m_q.nodes[pChildren[i]];
			if ( tChild.n_words!=1 || tChild.n_children )
				return Fail ( MGPU_E_UNSUPPORTED );
		}
		if ( ( tNode.op==MGPU_OP_NEAR && nChildren>2 ) || nChildren>MAX_PHRASE_WORDS || d.m_nNWay>=MAX_NWAY )
			return Fail ( MGPU_E_UNSUPPORTED );
		PNode_t t;
		t.m_eKind = PN_NWAY;
		t.m_iNWay = d.m_nNWay;
		DevNWay_t & n = d.m_dNWay[d.m_nNWay++];
		n.m_eKind = tNode.op==MGPU_OP_NEAR ? NWAY_NEAR : tNode.op==MGPU_OP_BEFORE ? NWAY_BEFORE : NWAY_NOTNEAR;
		n.m_iOpArg = tNode.oparg;
		n.m_nWords = nChildren;
		for ( int i=0; i<nChildren; ++i )
		{
			const mgpu_xqnode & tChild = m_q.nodes[pChildren[i]];
			int iLeaf = AddLeaf ( tChild, tChild.first_word, i );
			if ( iLeaf<0 )
				return -1;
			t.m_dLeaves.push_back ( iLeaf );
			n.m_dLeaf[i] = iLeaf;
			n.m_dAtomPos[i] = m_dLeaves[iLeaf].m_iAtomPos;
			n.m_dCount[i] = 1;
		}
		if ( n.m_eKind==NWAY_BEFORE )
			t.m_dRegOrder = t.m_dLeaves;	// ExtOrder_c::GetQwords walks the children as written
		if ( n.m_eKind!=NWAY_NOTNEAR )
			RefSort ( t.m_dLeaves, [this] ( int a, int b ) { return m_dLeaves[a].Docs()<m_dLeaves[b].Docs(); } );	// the AND chain under the acceptor
		return NewNode ( t );
	}

	/// ExtQuorum_c (src/searchnode.cpp:4319-4650): keywords the query repeats fold into the first occurrence with a count (ctor :4342-4404),
	/// the children are kept by query position; a document matches when the counts of the keywords on it (each up to its hits there)
	/// reach the threshold. One quorum node per query on the GPU path.
	int CreateQuorumNode ( const mgpu_xqnode & tNode )
	{
		DevQuery_t & d = m_tFull;
		if ( d.m_nNWay>=MAX_NWAY )
			return Fail ( MGPU_E_UNSUPPORTED );
		for ( int j=0; j<d.m_nNWay; ++j )
			if ( d.m_dNWay[j].m_eKind==NWAY_QUORUM )
				return Fail ( MGPU_E_UNSUPPORTED );
		struct Child_t { int m_iWord; int m_iCount; int m_iNodePos; };
		std::vector<Child_t> dChildren;
		for ( int i=0; i<tNode.n_words; ++i )
		{
			const int iWord = tNode.first_word+i;
			if ( !m_q.words[iWord].word )
				return Fail ( MGPU_E_BAD_QUERY );
			size_t iParent = dChildren.size();
			for ( size_t k=0; k<dChildren.size(); ++k )
				if ( !strcmp ( m_q.words[dChildren[k].m_iWord].word, m_q.words[iWord].word ) )
					iParent = k;
			if ( iParent<dChildren.size() )
				dChildren[iParent].m_iCount++;
			else
				dChildren.push_back ( { iWord, 1, i } );
		}
		std::sort ( dChildren.begin(), dChildren.end(), [this] ( const Child_t & a, const Child_t & b ) { return m_q.words[a.m_iWord].atom_pos<m_q.words[b.m_iWord].atom_pos; } );
		if ( (int)dChildren.size()>MAX_PHRASE_WORDS )
			return Fail ( MGPU_E_UNSUPPORTED );
		PNode_t t;
		t.m_eKind = PN_NWAY;
		t.m_iNWay = d.m_nNWay;
		DevNWay_t & n = d.m_dNWay[d.m_nNWay++];
		n.m_eKind = NWAY_QUORUM;
		n.m_iOpArg = tNode.oparg;
		n.m_nWords = (int)dChildren.size();
		for ( size_t i=0; i<dChildren.size(); ++i )
		{
			if ( dChildren[i].m_iCount>255 )
				return Fail ( MGPU_E_UNSUPPORTED );
			int iLeaf = AddLeaf ( tNode, dChildren[i].m_iWord, dChildren[i].m_iNodePos );
			if ( iLeaf<0 )
				return -1;
			t.m_dLeaves.push_back ( iLeaf );
			n.m_dLeaf[i] = iLeaf;
			n.m_dAtomPos[i] = m_dLeaves[iLeaf].m_iAtomPos;
			n.m_dCount[i] = (uint8_t)dChildren[i].m_iCount;
		}
		return NewNode ( t );
	}

	/// ExtNode_i::GetQwords in eval-tree order (src/searchnode.cpp:2030-2057, 3244-3253, ExtTwofer_c)
	void GetQwords ( int iNode )
	{
		const PNode_t & t = m_dNodes[iNode];
		switch ( t.m_eKind )
		{
		case PN_TERM:		Register ( t.m_iLeaf ); break;
		case PN_MULTIAND:
		case PN_MULTIOR:
		case PN_NWAY:		for ( int l : ( t.m_dRegOrder.empty() ? t.m_dLeaves : t.m_dRegOrder ) ) Register ( l ); break;
		default:			GetQwords ( t.m_iLeft ); GetQwords ( t.m_iRight ); break;
		}
	}

	void Register ( int iLeaf )
	{
		PLeaf_t & l = m_dLeaves[iLeaf];
		const mgpu_xqkeyword & w = m_q.words[l.m_iWord];
		auto it = m_hQwords.find ( w.word );
		if ( l.m_bNotWeighted || it!=m_hQwords.end() )
			return;
		l.m_bOwnsIDF = true;
		m_hQwords.emplace ( w.word, Qword_t { l.LocalDocs(), w.boost, l.m_iWord, 0.0f } );
	}

	/// emits the tile program; returns the alive value of v[iSp]
	int Emit ( int iNode, int iSp )
	{
		if ( iSp>=MAX_STACK )
		{
			Fail ( MGPU_E_UNSUPPORTED );
			return 1;
		}
		m_iMaxSp = std::max ( m_iMaxSp, iSp );
		const PNode_t & t = m_dNodes[iNode];
		switch ( t.m_eKind )
		{
		case PN_TERM:
			AddOp ( OP_TERM_SET, iSp, 0, 0, 0, t.m_iLeaf, 1 );
			return 1;
		case PN_MULTIOR:
			{
				AddOp ( OP_TERM_SET, iSp, 0, 0, 0, t.m_dLeaves[0], 1 );
				for ( size_t i=1; i<t.m_dLeaves.size(); ++i )
					AddOp ( OP_TERM_OR, iSp, 0, 1, 0, t.m_dLeaves[i], 1 );
				return 1;
			}
		case PN_MULTIAND:
		case PN_NWAY:
			{
				const int eNWay = t.m_eKind==PN_NWAY ? m_tFull.m_dNWay[t.m_iNWay].m_eKind : -1;
				AddOp ( OP_TERM_SET, iSp, 0, 0, 0, t.m_dLeaves[0], 1 );
				int iAlive = 1;
				if ( eNWay==NWAY_NOTNEAR )
					AddOp ( OP_TERM_MAYBE, iSp, 0, 1, 0, t.m_dLeaves[1], 1 );	// every MUST document is a candidate; the NOT keyword only brings its hits
				else if ( eNWay==NWAY_QUORUM )
					for ( size_t i=1; i<t.m_dLeaves.size(); ++i )
						AddOp ( OP_TERM_OR, iSp, 0, 1, 0, t.m_dLeaves[i], 1 );	// candidates = documents holding any of the keywords
				else
					for ( size_t i=1; i<t.m_dLeaves.size(); ++i, ++iAlive )
						AddOp ( OP_TERM_AND, iSp, 0, iAlive, 0, t.m_dLeaves[i], iAlive+1 );
				if ( t.m_eKind==PN_NWAY )
					AddOp ( OP_NWAY, iSp, 0, iAlive, 0, 0, iAlive, t.m_iNWay );
				return iAlive;
			}
		default:
			{
				int iAliveL = Emit ( t.m_iLeft, iSp );
				const PNode_t & r = m_dNodes[t.m_iRight];
				if ( r.m_eKind==PN_TERM )
				{
					switch ( t.m_eKind )
					{
					case PN_AND:	AddOp ( OP_TERM_AND, iSp, 0, iAliveL, 0, r.m_iLeaf, iAliveL+1 ); return iAliveL+1;
					case PN_OR:		AddOp ( OP_TERM_OR, iSp, 0, iAliveL, 0, r.m_iLeaf, iAliveL ); return iAliveL;
					case PN_MAYBE:	AddOp ( OP_TERM_MAYBE, iSp, 0, iAliveL, 0, r.m_iLeaf, iAliveL ); return iAliveL;
					default:		AddOp ( OP_TERM_ANDNOT, iSp, 0, iAliveL, 0, r.m_iLeaf, iAliveL ); return iAliveL;
					}
				}
				int iAliveR = Emit ( t.m_iRight, iSp+1 );
				int eCode = t.m_eKind==PN_AND ? OP_VEC_AND : t.m_eKind==PN_OR ? OP_VEC_OR : t.m_eKind==PN_MAYBE ? OP_VEC_MAYBE : OP_VEC_ANDNOT;
				AddOp ( eCode, iSp, iSp+1, iAliveL, iAliveR, 0, iAliveL );
				return iAliveL;
			}
		}
	}

	void AddOp ( int eCode, int iDst, int iSrc, int iAliveDst, int iAliveSrc, int iLeaf, int iAliveOut, int iArg=0 )
	{
		if ( m_tFull.m_nOps>=MAX_OPS || iAliveOut>250 )
		{
			Fail ( MGPU_E_UNSUPPORTED );
			return;
		}
		DevOp_t & o = m_tFull.m_dOps[m_tFull.m_nOps++];
		o.m_eCode = (uint8_t)eCode;
		o.m_uDst = (uint8_t)iDst;
		o.m_uSrc = (uint8_t)iSrc;
		o.m_uAliveDst = (uint8_t)iAliveDst;
		o.m_uAliveSrc = (uint8_t)iAliveSrc;
		o.m_uLeaf = (uint8_t)iLeaf;
		o.m_uArg = (uint8_t)iArg;
		o.m_uAliveOut = (uint8_t)iAliveOut;
	}

	int Run()
	{
		DevQuery_t & d = m_tFull;
		memset ( &d, 0, sizeof(d) );
		m_tOut.m_dLeafTerms.m_n = m_tOut.m_dLeafWord.m_n = 0;
		m_tOut.m_dWordStats.assign ( std::max ( m_q.n_words, 0 ), mgpu_wordstat { 0, 0 } );
		if ( m_q.n_words>MAX_LEAVES )
			return MGPU_E_UNSUPPORTED;	// (more keywords than a device query has leaves)
		for ( int i=0; i<m_q.n_words && m_q.words; ++i )
			if ( m_q.words[i].word )
				if ( const TermInfo_t * p = m_tIndex.FindTerm ( m_q.words[i].word ) )
				{
					m_tOut.m_dWordStats[i].docs = p->m_iDocs;
					m_tOut.m_dWordStats[i].hits = p->m_iHits;
				}

		if ( m_q.ranker<MGPU_RANK_PROXIMITY_BM25 || m_q.ranker>MGPU_RANK_SPH04 )
			return MGPU_E_UNSUPPORTED;	// EXPR / EXPORT / PLUGIN rankers need the expression engine
		// counts without arrays
		if ( ( m_q.n_nodes>0 && !m_q.nodes ) || ( m_q.n_children>0 && !m_q.children ) || ( m_q.n_words>0 && !m_q.words )
			|| ( m_q.n_filters>0 && !m_q.filters ) || ( m_q.n_sort_keys>0 && !m_q.sort_keys ) || ( m_q.n_field_weights>0 && !m_q.field_weights )
			|| m_q.n_nodes<0 || m_q.n_children<0 || m_q.n_words<0 || m_q.n_filters<0 || m_q.n_sort_keys<0 || m_q.n_field_weights<0 )
			return MGPU_E_BAD_QUERY;

		int iRoot = -1;
		if ( m_q.n_nodes>0 && m_q.root>=0 )
			iRoot = Create ( m_q.root );
		if ( m_iError!=MGPU_OK )
			return m_iError;

		// ranker selection, sphCreateRanker src/sphinxsearch.cpp:4192-4226
		bool bSingleWord = false;
		if ( m_q.n_nodes>0 && m_q.root>=0 && m_q.root<m_q.n_nodes )
			bSingleWord = ( m_q.nodes[m_q.root].n_words==1 && m_q.nodes[m_q.root].n_children==0 );
		int eRanker = m_q.ranker;
		if ( eRanker==MGPU_RANK_PROXIMITY_BM25 && bSingleWord )
			eRanker = MGPU_RANK_BM25;	// ExtRanker_WeightSum_c<WITH_BM25>
		d.m_eRanker = eRanker;
		// ExtRanker_State_T (src/sphinxsearch.cpp:4189-4232); a single keyword under PROXIMITY / PROXIMITY_BM25 gets ExtRanker_WeightSum_c
		d.m_bStateRanker = ( eRanker==MGPU_RANK_PROXIMITY_BM25 || eRanker==MGPU_RANK_WORDCOUNT || eRanker==MGPU_RANK_MATCHANY
			|| eRanker==MGPU_RANK_FIELDMASK || eRanker==MGPU_RANK_SPH04 || ( eRanker==MGPU_RANK_PROXIMITY && !bSingleWord ) ) ? 1 : 0;
		d.m_bNeedHits = ( d.m_bStateRanker || d.m_nNWay>0 || m_bAnyTermPos ) ? 1 : 0;	// position filters look at the hits too
		const bool bUseBM25 = ( eRanker==MGPU_RANK_BM25 || eRanker==MGPU_RANK_PROXIMITY_BM25 || eRanker==MGPU_RANK_SPH04 );
		{
			// HasQwordDupes, src/sphinxsearch.cpp:4148-4164
			std::unordered_map<std::string,int> hSeen;
			for ( int i=0; i<m_q.n_words; ++i )
				if ( m_q.words[i].word && !hSeen.emplace ( m_q.words[i].word, 1 ).second )
					d.m_bDupes = 1;
		}

		if ( iRoot>=0 )
		{
			GetQwords ( iRoot );
			// IDF, src/sphinxsearch.cpp:4293-4361
			const int iQwords = (int)m_hQwords.size();
			d.m_nQwords = iQwords;
			d.m_iMaxQpos = -1;	// max in-query position over the non-excluded keywords (ExtTerm_T::GetQwords return value)
			for ( const PLeaf_t & l : m_dLeaves )
				d.m_iMaxQpos = std::max ( d.m_iMaxQpos, m_q.words[l.m_iWord].excluded ? -1 : l.m_iAtomPos );
			const int64_t iTotalDocuments = m_q.total_docs>0 ? m_q.total_docs : (int64_t)m_tIndex.m_tHdr.m_iTotalDocuments;
			for ( auto & kv : m_hQwords )
			{
				Qword_t & w = kv.second;
				int64_t iTermDocs = w.m_iDocs;
				if ( m_q.word_docs && m_q.word_docs[w.m_iFirstWord]>=0 )
					iTermDocs = m_q.word_docs[w.m_iFirstWord];
				float fIDF = 0.0f;
				if ( iTermDocs )
				{
					const int64_t iTotalClamped = std::max ( iTotalDocuments, iTermDocs );
					float fLogTotal = logf ( float ( 1+iTotalClamped ) );
					if ( !m_q.plain_idf )
						fIDF = logf ( float ( iTotalClamped-iTermDocs+1 ) / float ( iTermDocs ) ) / ( 2*fLogTotal );
					else
						fIDF = logf ( float ( iTotalClamped ) / float ( iTermDocs ) ) / ( 2*fLogTotal );
				}
				if ( !m_q.unnormalized_tfidf )
					fIDF /= iQwords;
				w.m_fIDF = fIDF * w.m_fBoost;
			}

			int iAlive = Emit ( iRoot, 0 );
			if ( m_iError!=MGPU_OK )
				return m_iError;
			d.m_uAliveRoot = iAlive;
		}

		// AND chains: where to resume when a tile has no candidate left (consecutive TERM_ANDs into the same level, plus
		// the acceptor of a phrase/proximity node; an empty vector stays empty through all of them)
		for ( int i=0; i<d.m_nOps; ++i )
			if ( d.m_dOps[i].m_eCode==OP_TERM_AND || d.m_dOps[i].m_eCode==OP_TERM_SET )
			{
				int j = i+1;
				while ( j<d.m_nOps && d.m_dOps[j].m_eCode==OP_TERM_AND && d.m_dOps[j].m_uDst==d.m_dOps[i].m_uDst )
					++j;
				if ( j<d.m_nOps && d.m_dOps[j].m_eCode==OP_NWAY && d.m_dOps[j].m_uDst==d.m_dOps[i].m_uDst )
					++j;
				d.m_dOps[i].m_uSrc = (uint8_t)( 1+j );
			}

		if ( !m_tIndex.m_tOpt.m_bChain )
			for ( int i=0; i<d.m_nOps; ++i )
				if ( d.m_dOps[i].m_eCode<=OP_TERM_MAYBE )
					d.m_dOps[i].m_uSrc = 0;
		for ( int i=0; i<d.m_nOps; ++i )
			if ( d.m_dOps[i].m_eCode<=OP_TERM_MAYBE && !( d.m_dOps[i].m_eCode==OP_TERM_AND && d.m_dOps[i].m_uSrc ) )
				d.m_uPreMask |= 1u<<d.m_dOps[i].m_uLeaf;
		for ( int i=0; i<d.m_nOps; ++i )
			if ( d.m_dOps[i].m_eCode==OP_TERM_SET || d.m_dOps[i].m_eCode==OP_TERM_OR )
				d.m_uOrigMask |= 1u<<d.m_dOps[i].m_uLeaf;
		d.m_bPureOr = ( m_iMaxSp==0 && d.m_nOps>0 && m_tIndex.m_tOpt.m_bRegOr ) ? 1 : 0;
		for ( int i=0; i<d.m_nOps && d.m_bPureOr; ++i )
			d.m_bPureOr = ( d.m_dOps[i].m_eCode==( i ? OP_TERM_OR : OP_TERM_SET ) && d.m_dOps[i].m_uDst==0 ) ? 1 : 0;
		d.m_iDriverLeaf = -1;
		if ( m_tIndex.m_tOpt.m_bJump && d.m_nOps>0 && d.m_dOps[0].m_eCode==OP_TERM_SET && (int)d.m_dOps[0].m_uSrc==d.m_nOps+1 )
			d.m_iDriverLeaf = d.m_dOps[0].m_uLeaf;

		// DNF shape: [SET AND*] then ( [SET AND*] on level 1 + VEC_OR(0,1) | TERM_OR on level 0 )*  -> groups for the intersection kernel
		{
			int i = 0, nGroups = 0;
			bool bOk = d.m_nOps>0;
			auto fnGroup = [&] ( int iDst ) -> bool
			{
				if ( i>=d.m_nOps || d.m_dOps[i].m_eCode!=OP_TERM_SET || d.m_dOps[i].m_uDst!=iDst || nGroups>=MAX_GROUPS )
					return false;
				int j = i+1;
				while ( j<d.m_nOps && d.m_dOps[j].m_eCode==OP_TERM_AND && d.m_dOps[j].m_uDst==iDst )
					++j;
				// `a b -c -d`: negated keywords behind the group, if they end the whole program (one group; and_kernel probes and rejects)
				if ( i==0 && iDst==0 && j<d.m_nOps && d.m_dOps[j].m_eCode==OP_TERM_ANDNOT && m_tIndex.m_tOpt.m_bGroupNeg )
				{
					int k = j;
					while ( k<d.m_nOps && d.m_dOps[k].m_eCode==OP_TERM_ANDNOT && d.m_dOps[k].m_uDst==0 )
						++k;
					if ( k==d.m_nOps )
					{
						d.m_bGroupNeg = 1;
						j = k;
					}
				}
				d.m_dGroupOp0[nGroups] = (uint8_t)i;
				d.m_dGroupOps[nGroups] = (uint8_t)( j-i );
				++nGroups;
				i = j;
				return true;
			};
			bOk = bOk && fnGroup ( 0 );
			bool bChainNWay = false;
			if ( bOk && i==d.m_nOps-1 && d.m_dOps[i].m_eCode==OP_NWAY && d.m_dOps[i].m_uDst==0 )
			{
				bChainNWay = true;	// a phrase / proximity node alone: AND chain + acceptor
				++i;
			}
			while ( bOk && i<d.m_nOps )
			{
				if ( d.m_dOps[i].m_eCode==OP_TERM_OR && d.m_dOps[i].m_uDst==0 && nGroups<MAX_GROUPS )
				{
					d.m_dGroupOp0[nGroups] = (uint8_t)i;
					d.m_dGroupOps[nGroups] = 1;
					++nGroups;
					++i;
				} else if ( fnGroup ( 1 ) && i<d.m_nOps && d.m_dOps[i].m_eCode==OP_VEC_OR && d.m_dOps[i].m_uDst==0 && d.m_dOps[i].m_uSrc==1 )
					++i;
				else
					bOk = false;
			}
			bool bAnyMulti = false;
			for ( int g=0; g<nGroups; ++g )
				bAnyMulti |= d.m_dGroupOps[g]>1;
			if ( d.m_bGroupNeg && ( !bOk || nGroups!=1 || d.m_bNeedHits ) )
			{
				// (hit-consuming programs: the right side of an ANDNOT never emits hits; they stay on dense tiles)
				bOk = false;
				d.m_bGroupNeg = 0;
			}
			d.m_nGroups = ( bOk && bAnyMulti && m_tIndex.m_tOpt.m_bDnf ) ? nGroups : ( d.m_iDriverLeaf>=0 ? 1 : 0 );
			if ( d.m_nGroups==1 && d.m_iDriverLeaf>=0 )
			{
				d.m_dGroupOp0[0] = 0;
				d.m_dGroupOps[0] = (uint8_t)( d.m_nOps - ( bChainNWay ? 1 : 0 ) );
			}
			bool bTileOnly = false;
			for ( int j=0; j<d.m_nNWay; ++j )
				bTileOnly |= d.m_dNWay[j].m_eKind>=NWAY_BEFORE;	// these nodes rebuild the document's TF*IDF from per-slot records of eval_kernel
			if ( ( d.m_bNeedHits && d.m_nGroups>1 ) || m_bAnyTermPos || bTileOnly )
				d.m_nGroups = 0;	// hit-consuming DNF and position-filtered keywords stay on dense tiles
		}

		d.m_nLeaves = (int)m_dLeaves.size();
		const unsigned nIndexFields = (unsigned)m_tIndex.m_tHdr.m_dFields.size();
		m_tOut.m_dLeafTerms.assign ( m_dLeaves.size(), nullptr );
		m_tOut.m_dLeafWord.assign ( m_dLeaves.size(), 0 );
		for ( size_t i=0; i<m_dLeaves.size(); ++i )
		{
			const PLeaf_t & l = m_dLeaves[i];
			DevLeaf_t & t = d.m_dLeaves[i];
			t.m_iHot = -1;
			m_tOut.m_dLeafTerms[i] = l.m_pTerm;
			m_tOut.m_dLeafWord[i] = l.m_iWord;
			if ( l.m_pTerm )
			{
				t.m_uFirstBlk = l.m_pTerm->m_uFirstBlk;
				t.m_nBlocks = l.m_pTerm->m_nBlocks;
				t.m_nDocs = (uint32_t)l.m_pTerm->m_iDocs;
				t.m_uDoclistEnd = (uint64_t)( l.m_pTerm->m_iDoclistOffset+l.m_pTerm->m_iDoclistLength-1 );
				m_tOut.m_iCost += l.m_pTerm->m_iDocs;
				m_tOut.m_iAlgBytes += l.m_pTerm->m_iDoclistLength + l.m_pTerm->m_iSkiplistBytes;
			}
			// fields the index does not have never match; keeping the mask tight lets the dense store use the spare field bits
			t.m_uQueriedFields = l.m_uFields & ( nIndexFields>=32 ? 0xFFFFFFFFu : ( ( 1u<<nIndexFields )-1u ) );
			t.m_iTermPos = l.m_iTermPos;
			t.m_fIDF = 0.0f;
			if ( bUseBM25 && l.m_bOwnsIDF )
				t.m_fIDF = m_hQwords.at ( m_q.words[l.m_iWord].word ).m_fIDF;
			t.m_uAtomPos = (uint16_t)l.m_iAtomPos;
			t.m_uNodePos = (uint16_t)l.m_iNodePos;
		}
		m_tOut.m_nStack = m_iMaxSp+1;

		// field weights (already bound by the caller: BindWeights, src/sphinx.cpp:13903-13943)
		const int nFields = (int)m_tIndex.m_tHdr.m_dFields.size();
		d.m_nWeights = std::min ( nFields, 32 );
		for ( int i=0; i<MAX_FIELDS; ++i )
			d.m_dWeights[i] = ( m_q.field_weights && i<m_q.n_field_weights ) ? m_q.field_weights[i] : 1;

		// filters
		if ( m_q.n_filters>MAX_FILTERS )
			return MGPU_E_UNSUPPORTED;
		d.m_nFilters = std::max ( m_q.n_filters, 0 );
		for ( int i=0; i<d.m_nFilters; ++i )
		{
			const mgpu_filter & f = m_q.filters[i];
			if ( f.attr<0 || f.attr>=(int)m_tIndex.m_tHdr.m_dAttrs.size() )
				return MGPU_E_BAD_QUERY;
			const SchemaAttr_t & a = m_tIndex.m_tHdr.m_dAttrs[f.attr];
			if ( ( a.m_iBitCount!=32 && a.m_iBitCount!=64 ) || ( a.m_iBitOffset & 31 ) )
				return MGPU_E_UNSUPPORTED;
			DevFilter_t & t = d.m_dFilters[i];
			t.m_eKind = f.kind;
			t.m_iMin = f.min_value;
			t.m_iMax = f.max_value;
			t.m_bExclude = f.exclude;
			t.m_iDwordOff = (int)( a.m_iBitOffset/32 );
			t.m_iBitCount = (int)a.m_iBitCount;
			if ( f.kind==MGPU_FILTER_VALUES )
			{
				if ( f.n_values<0 || ( f.n_values>0 && !f.values ) )
					return MGPU_E_BAD_QUERY;
				if ( f.n_values>MAX_FILTER_VALUES )
					return MGPU_E_UNSUPPORTED;
				t.m_nValues = f.n_values;
				for ( int k=0; k<f.n_values; ++k )
					t.m_dValues[k] = f.values[k];
			} else if ( f.kind!=MGPU_FILTER_RANGE )
				return MGPU_E_UNSUPPORTED;
		}

		// sort keys -> 64-bit packed key, most significant first
		if ( m_q.n_sort_keys>5 )
			return MGPU_E_BAD_QUERY;
		d.m_nSortKeys = std::max ( m_q.n_sort_keys, 0 );
		int iCur = 64;
		for ( int i=0; i<d.m_nSortKeys; ++i )
		{
			const mgpu_sortkey & k = m_q.sort_keys[i];
			DevSortKey_t & t = d.m_dSortKeys[i];
			t.m_eKind = k.kind;
			t.m_bDesc = k.desc;
			int iBits = 32;
			if ( k.kind==MGPU_KEYPART_INT )
			{
				if ( k.attr<0 || k.attr>=(int)m_tIndex.m_tHdr.m_dAttrs.size() )
					return MGPU_E_BAD_QUERY;
				const SchemaAttr_t & a = m_tIndex.m_tHdr.m_dAttrs[k.attr];
				if ( ( a.m_iBitCount!=32 && a.m_iBitCount!=64 ) || ( a.m_iBitOffset & 31 ) )
					return MGPU_E_UNSUPPORTED;
				t.m_iDwordOff = (int)( a.m_iBitOffset/32 );
				t.m_iBitCount = (int)a.m_iBitCount;
				iBits = t.m_iBitCount;
			} else if ( k.kind==MGPU_KEYPART_FLOAT )
			{
				if ( k.attr<0 || k.attr>=(int)m_tIndex.m_tHdr.m_dAttrs.size() )
					return MGPU_E_BAD_QUERY;
				const SchemaAttr_t & a = m_tIndex.m_tHdr.m_dAttrs[k.attr];
				if ( a.m_iBitCount!=32 || ( a.m_iBitOffset & 31 ) )
					return MGPU_E_UNSUPPORTED;
				t.m_iDwordOff = (int)( a.m_iBitOffset/32 );
				t.m_iBitCount = 32;
			} else if ( k.kind!=MGPU_KEYPART_WEIGHT && k.kind!=MGPU_KEYPART_ROWID )
				return MGPU_E_UNSUPPORTED;
			iCur -= iBits;
			if ( iCur<0 )
				return MGPU_E_UNSUPPORTED;	// more than 64 bits of sort keys
			t.m_iShift = iCur;
			if ( k.kind==MGPU_KEYPART_INT && m_tOut.m_iFirstIntKeyShift<0 )
			{
				m_tOut.m_iFirstIntKeyShift = iCur;
				m_tOut.m_iFirstIntKeyBits = iBits;
				m_tOut.m_bFirstIntKeyDesc = k.desc!=0;
			}
		}

		d.m_iIndexWeight = m_q.index_weight ? m_q.index_weight : 1;
		d.m_iMaxMatches = m_q.max_matches>0 ? m_q.max_matches : 1000;
		if ( d.m_iMaxMatches>65536 )
			return MGPU_E_UNSUPPORTED;
		return MGPU_OK;
	}
};

} // namespace


int PlanQuery ( const Index_c & tIndex, const mgpu_query & tQuery, PlannedQuery_t & tOut, DevQueryExt_t & tExt, bool & bHasExt )
{
	Planner_c tPlanner ( tIndex, tQuery, tOut );
	tOut.m_iStatus = tPlanner.Run();
	DevQuery_t & tFull = tPlanner.m_tFull;
	for ( int k=0; k<tFull.m_nSortKeys; ++k )
		tFull.m_bWeightKey |= tFull.m_dSortKeys[k].m_eKind==1;
	tFull.m_iExt = -1;
	tOut.m_tDev = static_cast<const DevQueryCore_t &>( tFull );
	bHasExt = tOut.m_iStatus==MGPU_OK && ( tFull.m_nFilters>0 || tFull.m_nSortKeys>0 || tFull.m_nNWay>0 );
	if ( bHasExt )
		tExt = static_cast<const DevQueryExt_t &>( tFull );
	return tOut.m_iStatus;
}

/// A plan made on one rowid-range shard, re-bound to another shard of the same index: the program, IDFs (global statistics), weights,
/// filters and sort keys do not depend on the shard; the keywords' dictionary entries and the statistics derived from them do.
/// Only valid for queries that carry global word_docs (so that no keyword order was decided from shard-local counts).
void RebindPlan ( const Index_c & tIndex, const mgpu_query & tQuery, PlannedQuery_t & tPlan, const int32_t * pWordIds, const TermInfo_t * const * pTermOfId )
{
	// (the sharded handle reports the keywords' statistics from its global table: the shards' own are not needed then)
	for ( size_t i=0; i<tPlan.m_dWordStats.size() && !pWordIds; ++i )
	{
		tPlan.m_dWordStats[i] = mgpu_wordstat { 0, 0 };
		if ( tQuery.words && tQuery.words[i].word )
			if ( const TermInfo_t * p = tIndex.FindTerm ( tQuery.words[i].word ) )
			{
				tPlan.m_dWordStats[i].docs = p->m_iDocs;
				tPlan.m_dWordStats[i].hits = p->m_iHits;
			}
	}
	if ( tPlan.m_iStatus!=MGPU_OK )
		return;
	tPlan.m_iCost = 0;
	tPlan.m_iAlgBytes = 0;
	for ( size_t l=0; l<tPlan.m_dLeafTerms.size(); ++l )
	{
		const int iWord = tPlan.m_dLeafWord[l];
		const TermInfo_t * pTerm = pWordIds ? ( pWordIds[iWord]>=0 ? pTermOfId[pWordIds[iWord]] : nullptr ) : tIndex.FindTerm ( tQuery.words[iWord].word );
		tPlan.m_dLeafTerms[l] = pTerm;
		DevLeaf_t & t = tPlan.m_tDev.m_dLeaves[l];
		t.m_uFirstBlk = pTerm ? pTerm->m_uFirstBlk : 0;
		t.m_nBlocks = pTerm ? pTerm->m_nBlocks : 0;
		t.m_nDocs = pTerm ? (uint32_t)pTerm->m_iDocs : 0;
		t.m_uDoclistEnd = pTerm ? (uint64_t)( pTerm->m_iDoclistOffset+pTerm->m_iDoclistLength-1 ) : 0;
		t.m_iHot = -1;
		t.m_uListOff = 0;
		if ( pTerm )
		{
			tPlan.m_iCost += pTerm->m_iDocs;
			tPlan.m_iAlgBytes += pTerm->m_iDoclistLength + pTerm->m_iSkiplistBytes;
		}
	}
}

//////////////////////////////////////////////////////////////////////////
// engine options
//////////////////////////////////////////////////////////////////////////

bool EngineOptions_t::Set ( const char * szName, int64_t iValue )
{
	struct Opt_t { const char * m_szName; int EngineOptions_t::* m_pField; int64_t m_iMin, m_iMax; };
	static const Opt_t dOpts[] =
	{
		{ "plan_threads",	&EngineOptions_t::m_iPlanThreads,	0, 256 },
		{ "timing",			&EngineOptions_t::m_bTiming,		0, 1 },
		{ "stats",			&EngineOptions_t::m_bStats,			0, 1 },
		{ "hot_store",		&EngineOptions_t::m_bHotStore,		0, 1 },
		{ "hot_div",		&EngineOptions_t::m_iHotDiv,		1, 1<<20 },
		{ "hot_min_uses",	&EngineOptions_t::m_iHotMinUses,	1, 1<<20 },
		{ "hot_gb",			&EngineOptions_t::m_iHotGB,			1, 160 },
		{ "or_range_tiles",	&EngineOptions_t::m_iOrRangeTiles,	1, 1<<20 },
		{ "dnf_pct",		&EngineOptions_t::m_iDnfPct,		1, 100 },
		{ "eager_hot",		&EngineOptions_t::m_bEagerHot,		0, 1 },
		{ "group_neg",		&EngineOptions_t::m_bGroupNeg,		0, 2 },
		{ "force_hot",		&EngineOptions_t::m_bForceHot,		0, 1 },
		{ "or_bits",		&EngineOptions_t::m_bOrBits,		0, 1 },
		{ "bits_dnf",		&EngineOptions_t::m_bBitsDnf,		0, 1 },
		{ "bits_dnf_div",	&EngineOptions_t::m_iBitsDnfDiv,	0, 1<<20 },
		{ "or_class",		&EngineOptions_t::m_bOrClass,		0, 1 },
		{ "dnf_class",		&EngineOptions_t::m_bDnfClass,		0, 1 },
		{ "and_kernel",		&EngineOptions_t::m_bAndKernel,		0, 1 },
		{ "dnf",			&EngineOptions_t::m_bDnf,			0, 1 },
		{ "chain",			&EngineOptions_t::m_bChain,			0, 1 },
		{ "reg_or",			&EngineOptions_t::m_bRegOr,			0, 1 },
		{ "jump",			&EngineOptions_t::m_bJump,			0, 1 },
	};
	for ( const Opt_t & t : dOpts )
		if ( !strcmp ( t.m_szName, szName ) )
		{
			if ( iValue<t.m_iMin || iValue>t.m_iMax )
				return false;
			this->*( t.m_pField ) = (int)iValue;
			return true;
		}
	return false;
}

//////////////////////////////////////////////////////////////////////////
// batch executor
//////////////////////////////////////////////////////////////////////////

Batch_c::~Batch_c()
{
	if ( m_pIndex )
	{
		cudaSetDevice ( m_pIndex->m_iDevice );
		m_dPlans.clear();
		if ( m_dPlans.capacity()>m_pIndex->m_dPlanCache.capacity() )
		{
			std::lock_guard<std::mutex> tGuard ( m_pIndex->m_tCacheLock );
			m_dPlans.swap ( m_pIndex->m_dPlanCache );
		}
	}
	// (a batch that never ran: its early K0 may still be reading the descriptors freed below, in this stream's order)
	if ( m_bHotPending && m_tEvHotDone )
		cudaStreamWaitEvent ( m_tStream, m_tEvHotDone, 0 );
	if ( m_tEvHotDone ) cudaEventDestroy ( m_tEvHotDone );
	if ( m_pIndex )
		--m_pIndex->m_nLiveBatches;
	if ( m_tEv0 ) cudaEventDestroy ( m_tEv0 );
	if ( m_tEv1 ) cudaEventDestroy ( m_tEv1 );
	if ( m_tEv2 ) cudaEventDestroy ( m_tEv2 );
	if ( m_tEvHot ) cudaEventDestroy ( m_tEvHot );
	for ( int c=0; c<NUM_CLASSES; ++c )
		if ( m_dEvClass[c] ) cudaEventDestroy ( m_dEvClass[c] );
}

static int Pow2Ceil ( int n )
{
	int p = 1;
	while ( p<n )
		p <<= 1;
	return p;
}

/// fn ( from, to, thread ) over [0, n) on nThreads host threads (contiguous ranges, thread t gets the t-th one); inline for one thread
template<typename FN>
static void ParallelFor ( int n, int nThreads, FN && fn )
{
	nThreads = std::max ( 1, std::min ( nThreads, n ) );
	if ( nThreads<=1 )
	{
		fn ( 0, n, 0 );
		return;
	}
	std::vector<std::thread> dThreads;
	for ( int t=1; t<nThreads; ++t )
		dThreads.emplace_back ( fn, (int)( (int64_t)n*t/nThreads ), (int)( (int64_t)n*( t+1 )/nThreads ), t );
	fn ( 0, (int)( (int64_t)n/nThreads ), 0 );
	for ( auto & t : dThreads )
		t.join();
}

int Batch_c::Prepare ( Index_c * pIndex, const mgpu_query * pQueries, int nQueries, const std::vector<PlannedQuery_t> * pTemplate, int nMaxThreads, bool bEagerHot,
	const int32_t * pWordIds, const size_t * pWordOff, const TermInfo_t * const * pTermOfId, const std::vector<DevQueryExt_t> * pTemplateExt )
{
	m_pIndex = pIndex;
	++pIndex->m_nLiveBatches;
	const EngineOptions_t tOpt = pIndex->m_tOpt;	// one consistent copy per batch
	CUDA_TRY ( cudaSetDevice ( pIndex->m_iDevice ), m_sError );
	const auto tStart = std::chrono::steady_clock::now();
	// the plan array of a 10k-query batch is 34 MB: taken over from the previous batch on this handle (mapped pages) instead of
	// faulting fresh ones in, single-threaded, on every call
	{
		std::lock_guard<std::mutex> tGuard ( pIndex->m_tCacheLock );
		m_dPlans.swap ( pIndex->m_dPlanCache );
	}
	m_dPlans.clear();
	m_dPlans.resize ( nQueries );
	if ( tOpt.m_bTiming )
		fprintf ( stderr, "[mgpu setup] %-24s %7.2f ms (%zu B per plan)\n", "plan array", std::chrono::duration<float,std::milli> ( std::chrono::steady_clock::now()-tStart ).count(), sizeof(PlannedQuery_t) );
	{
		// planning is per query and read-only on the index: spread big batches over the host cores
		int nThreads = (int)std::min<unsigned> ( std::max ( 1u, std::thread::hardware_concurrency() ), 16u );
		nThreads = std::max ( 1, std::min ( nThreads, nQueries/512 ) );
		if ( tOpt.m_iPlanThreads>0 )
			nThreads = tOpt.m_iPlanThreads;
		if ( nMaxThreads>0 )
			nThreads = std::min ( nThreads, nMaxThreads );
		// (extensions: collected per thread, numbered in query order afterwards)
		std::vector<std::vector<std::pair<int,DevQueryExt_t>>> dThreadExt ( std::max ( nThreads, 1 ) );
		auto fnPlan = [&] ( int iFrom, int iTo, int iThread )
		{
			DevQueryExt_t tExt;
			for ( int i=iFrom; i<iTo; ++i )
				if ( pTemplate )
				{
					// planned once for all shards (sharded.cpp): take the plan and bind its keywords to this shard's dictionary
					m_dPlans[i] = (*pTemplate)[i];
					RebindPlan ( *pIndex, pQueries[i], m_dPlans[i], pWordIds ? pWordIds+pWordOff[i] : nullptr, pTermOfId );
				} else
				{
					bool bHasExt = false;
					PlanQuery ( *pIndex, pQueries[i], m_dPlans[i], tExt, bHasExt );
					if ( bHasExt )
						dThreadExt[iThread].push_back ( { i, tExt } );
				}
		};
		ParallelFor ( nQueries, nThreads, fnPlan );
		for ( const auto & dMine : dThreadExt )
			for ( const auto & t : dMine )
			{
				m_dPlans[t.first].m_tDev.m_iExt = (int)m_dExt.size();
				m_dExt.push_back ( t.second );
			}
	}
	const auto tPlanned = std::chrono::steady_clock::now();
	const bool bTiming = tOpt.m_bTiming!=0;
	auto tMark = tPlanned;
	auto fnMark = [&] ( const char * sWhat )
	{
		if ( !bTiming )
			return;
		const auto tNow = std::chrono::steady_clock::now();
		fprintf ( stderr, "[mgpu setup] %-24s %7.2f ms\n", sWhat, std::chrono::duration<float,std::milli> ( tNow-tMark ).count() );
		tMark = tNow;
	};
	m_tStats.host_plan_ms = std::chrono::duration<float,std::milli> ( tPlanned-tStart ).count();

	// runnable queries; three launch classes (see engine.h). The plans are 3.4 KB apiece: every pass over a 10k-query batch walks
	// 34 MB, so the passes of this function run on a few host threads (contiguous query ranges, merged in thread order)
	int nSetupThreads = (int)std::min<unsigned> ( std::max ( 1u, std::thread::hardware_concurrency() ), 6u );
	nSetupThreads = std::max ( 1, std::min ( nSetupThreads, nQueries/1024 ) );
	if ( tOpt.m_iPlanThreads>0 )
		nSetupThreads = std::min ( nSetupThreads, tOpt.m_iPlanThreads );
	if ( nMaxThreads>0 )
		nSetupThreads = std::min ( nSetupThreads, nMaxThreads );
	std::vector<int> dDocOnly, dOrder[NUM_CLASSES];
	{
		struct Part_t { std::vector<int> m_dDocOnly, m_dHits1, m_dHits4; int m_iKMax = 1; int64_t m_iAlg = 0, m_iPostings = 0; };
		std::vector<Part_t> dParts ( nSetupThreads );
		ParallelFor ( nQueries, nSetupThreads, [&] ( int iFrom, int iTo, int t )
		{
			Part_t & tPart = dParts[t];
			for ( int i=iFrom; i<iTo; ++i )
				if ( m_dPlans[i].m_iStatus==MGPU_OK && m_dPlans[i].m_tDev.m_nOps>0 )
				{
					if ( m_dPlans[i].m_tDev.m_bNeedHits )
						( ( m_dPlans[i].m_tDev.m_nGroups==1 && tOpt.m_bAndKernel ) ? tPart.m_dHits4 : tPart.m_dHits1 ).push_back ( i );
					else
						tPart.m_dDocOnly.push_back ( i );
					tPart.m_iKMax = std::max ( tPart.m_iKMax, m_dPlans[i].m_tDev.m_iMaxMatches );
					tPart.m_iAlg += m_dPlans[i].m_iAlgBytes;
					tPart.m_iPostings += m_dPlans[i].m_iCost;
				}
		} );
		for ( const Part_t & tPart : dParts )
		{
			dDocOnly.insert ( dDocOnly.end(), tPart.m_dDocOnly.begin(), tPart.m_dDocOnly.end() );
			dOrder[1].insert ( dOrder[1].end(), tPart.m_dHits1.begin(), tPart.m_dHits1.end() );
			dOrder[4].insert ( dOrder[4].end(), tPart.m_dHits4.begin(), tPart.m_dHits4.end() );
			m_iKMax = std::max ( m_iKMax, tPart.m_iKMax );
			m_tStats.algorithmic_bytes += tPart.m_iAlg;
			m_tStats.postings += tPart.m_iPostings;
		}
	}
	if ( dDocOnly.empty() && dOrder[1].empty() && dOrder[4].empty() )
		return MGPU_OK;

	// the stream of this batch: captured here and used for its allocations, kernels, copies and frees (mgpu_index_set_stream
	// between prepare and run / free must not split them over two streams)
	m_tStream = pIndex->m_tStream;
	CUDA_TRY ( cudaEventCreate ( &m_tEv0 ), m_sError );
	CUDA_TRY ( cudaEventCreate ( &m_tEv1 ), m_sError );
	CUDA_TRY ( cudaEventCreate ( &m_tEv2 ), m_sError );
	CUDA_TRY ( cudaEventCreate ( &m_tEvHot ), m_sError );
	CUDA_TRY ( cudaEventCreate ( &m_tEvHotDone ), m_sError );
	for ( int c=0; c<NUM_CLASSES; ++c )
		CUDA_TRY ( cudaEventCreate ( &m_dEvClass[c] ), m_sError );

	m_dSlots.reserve ( nQueries );
	m_dDevToQuery.reserve ( nQueries );
	m_dItems.reserve ( (size_t)nQueries*4 );
	const uint32_t uRows = pIndex->m_tDev.m_uRows;
	const int nTiles = (int)( ( (uint64_t)uRows+TILE_W-1 )/TILE_W );
	int iMaxKeysPerQuery = 1;

	// hot keywords of the batch: shared by >= 2 doc-only queries and present in >= 1/iHotDiv of the rows -> dense store
	// (2 B/row read per query beats walking the compressed doclist from ~2 postings per 512-row mini-tile on)
	const int64_t iHotDiv = std::max ( 1, tOpt.m_iHotDiv );
	const int64_t iHotGB = std::max ( 1, tOpt.m_iHotGB );
	if ( pIndex->m_tHdr.m_dFields.size()<=8 && tOpt.m_bHotStore )
	{
		// (use counts and, below, store slots live in a flat per-keyword array of the index: 35k leaves of a 10k-query batch cost
		// 3 ms through hash maps, on the critical path in front of K0)
		std::vector<int32_t> & dUse = pIndex->m_dTermUse;
		std::vector<const TermInfo_t*> dTouched;
		// OR-of-AND-groups programs with a group led by a dense keyword cannot walk that driver block by block (and_kernel), and the
		// tile kernels pay ~10x a bitmap query for them: every keyword of theirs goes into the store, however rare or rarely used
		// (a sparse keyword costs K0 a few microseconds), so that the whole program runs on the bitmaps (orbits_kernel)
		const int32_t FORCE_BIT = 1<<30;
		const bool bForce = tOpt.m_bBitsDnf && tOpt.m_bOrBits && tOpt.m_bForceHot && pIndex->m_tHdr.m_dFields.size()<=4 && !pIndex->m_tDev.m_pDead && tOpt.m_bOrClass;
		{
			// (counters are bumped atomically; whoever moves one off zero lists the keyword)
			std::vector<std::vector<const TermInfo_t*>> dPartTouched ( nSetupThreads );
			ParallelFor ( (int)dDocOnly.size(), nSetupThreads, [&] ( int iFrom, int iTo, int t )
			{
				std::vector<const TermInfo_t*> & dMine = dPartTouched[t];
				for ( int k=iFrom; k<iTo; ++k )
				{
					const PlannedQuery_t & p = m_dPlans[dDocOnly[k]];
					for ( const TermInfo_t * pTerm : p.m_dLeafTerms )
						if ( pTerm && (int64_t)pTerm->m_iDocs*iHotDiv>=(int64_t)uRows && !__atomic_fetch_add ( &dUse[pTerm->m_iOrdinal], 1, __ATOMIC_RELAXED ) )
							dMine.push_back ( pTerm );
					const DevQueryCore_t & q = p.m_tDev;
					if ( !bForce || q.m_nGroups<=0 || q.m_bPureOr || q.m_eRanker!=1 || q.m_nFilters || q.m_nSortKeys || q.m_nWeights>4 )
						continue;
					bool bDense = false, bAll = true;
					for ( int g=0; g<q.m_nGroups; ++g )
					{
						const TermInfo_t * pDrv = p.m_dLeafTerms[q.m_dOps[q.m_dGroupOp0[g]].m_uLeaf];
						bDense |= pDrv && (int64_t)pDrv->m_iDocs*100>=(int64_t)uRows*std::max ( 1, tOpt.m_iDnfPct );
					}
					for ( const TermInfo_t * pTerm : p.m_dLeafTerms )
						bAll &= pTerm!=nullptr;
					if ( !bDense || !bAll )
						continue;
					for ( const TermInfo_t * pTerm : p.m_dLeafTerms )
						if ( !__atomic_fetch_or ( &dUse[pTerm->m_iOrdinal], FORCE_BIT, __ATOMIC_RELAXED ) )
							dMine.push_back ( pTerm );
				}
			} );
			for ( const auto & dMine : dPartTouched )
				dTouched.insert ( dTouched.end(), dMine.begin(), dMine.end() );
		}
		std::vector<std::pair<int64_t,const TermInfo_t*>> dHot;
		// ... or a keyword in >= 1/16 of the rows that only one query uses (small batches): one pass over its doclist into the
		// store and its bitmaps beats walking it posting by posting
		for ( const TermInfo_t * p : dTouched )
		{
			const int nUse = dUse[p->m_iOrdinal] & ( FORCE_BIT-1 );
			const bool bForced = ( dUse[p->m_iOrdinal] & FORCE_BIT )!=0;
			dUse[p->m_iOrdinal] = 0;
			if ( bForced )
				dHot.push_back ( { INT64_MAX/2, p } );
			else if ( nUse>=tOpt.m_iHotMinUses || (int64_t)p->m_iDocs*16>=(int64_t)uRows )
				dHot.push_back ( { (int64_t)nUse*p->m_iDocs, p } );
		}
		std::sort ( dHot.begin(), dHot.end(), [] ( const auto & a, const auto & b ) { return a.first>b.first || ( a.first==b.first && a.second->m_uFirstBlk<b.second->m_uFirstBlk ); } );
		m_iHotStride = (int64_t)nTiles*TILE_W;
		const size_t nMaxHot = std::min<size_t> ( 4096, ( (size_t)iHotGB<<30 )/( 2*(size_t)m_iHotStride ) );	// <= iHotGB of store
		if ( dHot.size()>nMaxHot )
			dHot.resize ( nMaxHot );
		int64_t iEscapeCap = 16;
		for ( const auto & t : dHot )
		{
			const TermInfo_t * p = t.second;
			dUse[p->m_iOrdinal] = (int)m_dHotTerms.size()+1;	// slot + 1
			DevLeaf_t tLeaf {};
			tLeaf.m_uFirstBlk = p->m_uFirstBlk;
			tLeaf.m_nBlocks = p->m_nBlocks;
			tLeaf.m_nDocs = (uint32_t)p->m_iDocs;
			tLeaf.m_uDoclistEnd = (uint64_t)( p->m_iDoclistOffset+p->m_iDoclistLength-1 );
			tLeaf.m_uQueriedFields = 0xFFFFFFFFu;
			tLeaf.m_iHot = -1;
			m_dHotTerms.push_back ( tLeaf );
			iEscapeCap += std::min<int64_t> ( p->m_iDocs, p->m_iHits/255 );	// at most hits/255 documents can hold >= 255 hits
		}
		m_iHotEscapeCap = (int)std::min<int64_t> ( iEscapeCap, 1<<26 );
		m_dHotBlkStart.assign ( 1, 0u );
		for ( const DevLeaf_t & t : m_dHotTerms )
			m_dHotBlkStart.push_back ( m_dHotBlkStart.back()+t.m_nBlocks );
		// indexes with <= 4 fields: per-field presence bitmaps next to the u16 rows (orbits_kernel)
		m_nHotBitFields = ( pIndex->m_tHdr.m_dFields.size()<=4 && !pIndex->m_tHdr.m_dFields.empty() ) ? (int)pIndex->m_tHdr.m_dFields.size() : 0;
		// keywords in >= 1/3 of the rows (any keyword whose idf can turn negative) also get two tf-level bitmaps
		m_dHotLvlSlot.assign ( m_dHotTerms.size(), -1 );
		if ( m_nHotBitFields )
			for ( size_t h=0; h<m_dHotTerms.size(); ++h )
				if ( (int64_t)m_dHotTerms[h].m_nDocs*3>=(int64_t)uRows )
					m_dHotLvlSlot[h] = m_nHotLvl++;
		if ( !m_dHotTerms.empty() )
			ParallelFor ( (int)dDocOnly.size(), nSetupThreads, [&] ( int iFrom, int iTo, int )
			{
				for ( int k=iFrom; k<iTo; ++k )
				{
					PlannedQuery_t & p = m_dPlans[dDocOnly[k]];
					for ( size_t l=0; l<p.m_dLeafTerms.size(); ++l )
					{
						const int iSlot = p.m_dLeafTerms[l] ? dUse[p.m_dLeafTerms[l]->m_iOrdinal]-1 : -1;
						if ( iSlot>=0 )
						{
							p.m_tDev.m_dLeaves[l].m_iHot = iSlot;
							if ( ( p.m_tDev.m_uOrigMask>>l ) & 1u )
								p.m_tDev.m_bOrigHot = 1;
							if ( p.m_tDev.m_iDriverLeaf==(int)l )
								p.m_tDev.m_iDriverLeaf = -1;
						}
					}
				}
			} );
		for ( const auto & t : dHot )
			dUse[t.second->m_iOrdinal] = 0;
	}

	// launch class 5 runs on the presence bitmaps when the store has them (every index class 5 admits has <= 4 fields)
	m_iOrMode = ( tOpt.m_bOrBits && m_nHotBitFields>0 && !m_dHotTerms.empty() ) ? 3 : 1;
	fnMark ( "hot keywords" );
	// pure AND queries led by a sparse keyword go to the intersection kernel, the rest of the doc-only ones to dense tiles
	const int64_t OR_RANGE_TILES = std::max ( 1, tOpt.m_iOrRangeTiles );	// 1024 x 2048 = 2M rows per item of class 5 (per-item costs outweigh finer ranges: 262144 rows was 25 % slower)
	const bool bNoAndKernel = !tOpt.m_bAndKernel;
	const bool bNoOrClass = !tOpt.m_bOrClass, bNoDnfClass = !tOpt.m_bDnfClass;	// (experiments)
	// a group's driver may sit in at most iDnfMul/iDnfDiv of the rows (option dnf_pct: percent, for experiments)
	const int64_t iDnfDiv = 100, iDnfMul = std::max ( 1, tOpt.m_iDnfPct );
	std::vector<uint8_t> dClassOf ( dDocOnly.size(), 0 );
	ParallelFor ( (int)dDocOnly.size(), nSetupThreads, [&] ( int iFrom, int iTo, int )
	{
	for ( int iDoc=iFrom; iDoc<iTo; ++iDoc )
	{
		// intersection kernel: DNF programs (1 group = pure AND) whose every group is led by a sparse keyword
		const int i = dDocOnly[iDoc];
		DevQueryCore_t & q = m_dPlans[i].m_tDev;
		bool bDnf = q.m_nGroups>0 && !bNoAndKernel;
		for ( int g=0; g<q.m_nGroups && bDnf; ++g )
		{
			// the group's driver (its rarest keyword) is walked block by block from the compressed doclist: fine unless it is dense
			const TermInfo_t * pDrv = m_dPlans[i].m_dLeafTerms[q.m_dOps[q.m_dGroupOp0[g]].m_uLeaf];
			bDnf = !pDrv || (int64_t)pDrv->m_iDocs*iDnfDiv<(int64_t)uRows*iDnfMul;
		}
		// pure OR programs under BM25, no dead rows, <= 4 fields, ordered by relevance or by attributes alone (a sort key on the
		// weight would need the bound inside the key): the bound + exact pass kernel; filters run inside its bound pass
		const bool bWeightKey = q.m_bWeightKey!=0;
		const bool bBoundBase = q.m_eRanker==1 && !bWeightKey && q.m_nWeights<=4 && !pIndex->m_tDev.m_pDead && !bNoOrClass;
		// OR-of-AND-groups programs (a pure AND is one group) whose multi-keyword groups hold hot keywords only intersect their
		// presence bitmaps on orbits_kernel instead of walking a driver's doclist block by block
		bool bBitsDnf = m_iOrMode==3 && tOpt.m_bBitsDnf && bBoundBase && !q.m_bPureOr && q.m_nGroups>0 && ( !q.m_bGroupNeg || tOpt.m_bGroupNeg>1 ) && !q.m_nFilters && !q.m_nSortKeys;
		for ( int g=0; g<q.m_nGroups && bBitsDnf; ++g )
			if ( q.m_dGroupOps[g]>1 )
			{
				// "bits_dnf_div" > 0 keeps groups whose rarest keyword sits in fewer than 1/div of the rows on and_kernel; measured on
				// the bench batch (div 16 / 64 / none: 103.2 / 100.4 / 98.7 ms per step) every hot group is better off on its bitmaps
				int64_t iMinDocs = INT64_MAX;
				for ( int iOp=q.m_dGroupOp0[g]; iOp<q.m_dGroupOp0[g]+q.m_dGroupOps[g]; ++iOp )
				{
					const DevLeaf_t & tLeaf = q.m_dLeaves[q.m_dOps[iOp].m_uLeaf];
					bBitsDnf = bBitsDnf && tLeaf.m_iHot>=0;
					iMinDocs = std::min<int64_t> ( iMinDocs, tLeaf.m_nDocs );
				}
				bBitsDnf = bBitsDnf && ( tOpt.m_iBitsDnfDiv<=0 || iMinDocs*tOpt.m_iBitsDnfDiv>=(int64_t)uRows );
			}
		if ( bBitsDnf )
		{
			dClassOf[iDoc] = 5;
			continue;
		}
		const bool bBoundOk = !bDnf && bBoundBase;
		const bool bOrClass = bBoundOk && q.m_bPureOr && !q.m_nFilters && !q.m_nSortKeys;	// the lean instantiation: relevance order, no filters
		// ... and the same passes with run-time options (class 6): pure OR programs with filters / attribute sort keys, and
		// OR-of-AND-groups programs (a dense driver kept them off the intersection kernel) whose multi-keyword groups hold at most
		// one keyword outside the dense store
		bool bHotDnf = bBoundOk && !bOrClass && ( q.m_bPureOr || q.m_nGroups>0 ) && !q.m_bGroupNeg && !bNoDnfClass;
		if ( bHotDnf && !q.m_bPureOr )
			for ( int g=0; g<q.m_nGroups && bHotDnf; ++g )
			{
				// a multi-keyword group may hold one sparse keyword (it then drives the group), the rest must be hot
				int nSparse = 0;
				for ( int iOp=q.m_dGroupOp0[g]; iOp<q.m_dGroupOp0[g]+q.m_dGroupOps[g]; ++iOp )
					nSparse += q.m_dLeaves[q.m_dOps[iOp].m_uLeaf].m_iHot<0 ? 1 : 0;
				bHotDnf = q.m_dGroupOps[g]==1 || nSparse<=1;
			}
		if ( ( !bDnf && !bHotDnf ) || ( bHotDnf && q.m_bPureOr ) )
			q.m_nGroups = 0;
		dClassOf[iDoc] = (uint8_t)( bDnf ? 2 : bOrClass ? 5 : bHotDnf ? 6 : m_dPlans[i].m_nStack>1 ? 3 : 0 );
	}
	} );
	for ( size_t iDoc=0; iDoc<dDocOnly.size(); ++iDoc )
		dOrder[dClassOf[iDoc]].push_back ( dDocOnly[iDoc] );

	// class 5 on orbits_kernel: every keyword outside the hot store is decoded once per run into a plain posting list
	if ( m_iOrMode==3 && !dOrder[5].empty() )
	{
		std::vector<int32_t> & dUse = pIndex->m_dTermUse;	// index into m_dListTerms + 1
		std::vector<const TermInfo_t*> dTouched;
		uint64_t uEntries = 0;
		for ( int i : dOrder[5] )
		{
			PlannedQuery_t & p = m_dPlans[i];
			for ( size_t l=0; l<p.m_dLeafTerms.size(); ++l )
			{
				const TermInfo_t * pTerm = p.m_dLeafTerms[l];
				if ( p.m_tDev.m_dLeaves[l].m_iHot>=0 || !pTerm )
					continue;
				int32_t & iList = dUse[pTerm->m_iOrdinal];
				if ( !iList )
				{
					DevLeaf_t tLeaf = p.m_tDev.m_dLeaves[l];
					tLeaf.m_uListOff = (uint32_t)std::min<uint64_t> ( uEntries, 0xFFFFFFFFu );
					m_dListTerms.push_back ( tLeaf );
					iList = (int32_t)m_dListTerms.size();
					dTouched.push_back ( pTerm );
					uEntries += ( (uint64_t)pTerm->m_nBlocks*32 + 31 ) & ~31ull;
				}
				p.m_tDev.m_dLeaves[l].m_uListOff = m_dListTerms[iList-1].m_uListOff;
			}
		}
		for ( const TermInfo_t * pTerm : dTouched )
			dUse[pTerm->m_iOrdinal] = 0;
		if ( uEntries>=( 1ull<<30 ) || uRows>=( 1u<<31 ) )
		{
			// (the lists would not fit / the candidate queue's flag bit is taken: the class stays on stream_kernel<512,1>)
			m_dListTerms.clear();
			m_iOrMode = 1;
		} else
		{
			m_nListEntries = (size_t)uEntries;
			m_dListBlkStart.assign ( 1, 0u );
			for ( const DevLeaf_t & t : m_dListTerms )
				m_dListBlkStart.push_back ( m_dListBlkStart.back()+t.m_nBlocks );
		}
	}

	// the store's descriptors go up now ...
	if ( !m_dHotTerms.empty() )
	{
		CUDA_TRY ( m_dHotDesc.AllocAsync ( m_dHotTerms.size(), m_tStream ), m_sError );
		CUDA_TRY ( cudaMemcpyAsync ( m_dHotDesc.m_p, m_dHotTerms.data(), m_dHotTerms.size()*sizeof(DevLeaf_t), cudaMemcpyHostToDevice, m_tStream ), m_sError );
		CUDA_TRY ( m_dHotLvlSlotDev.AllocAsync ( m_dHotLvlSlot.size(), m_tStream ), m_sError );
		CUDA_TRY ( cudaMemcpyAsync ( m_dHotLvlSlotDev.m_p, m_dHotLvlSlot.data(), m_dHotLvlSlot.size()*4, cudaMemcpyHostToDevice, m_tStream ), m_sError );
		CUDA_TRY ( m_dHotBlkStartDev.AllocAsync ( m_dHotBlkStart.size(), m_tStream ), m_sError );
		CUDA_TRY ( cudaMemcpyAsync ( m_dHotBlkStartDev.m_p, m_dHotBlkStart.data(), m_dHotBlkStart.size()*4, cudaMemcpyHostToDevice, m_tStream ), m_sError );
	}
	if ( !m_dListTerms.empty() )
	{
		CUDA_TRY ( m_dListDesc.AllocAsync ( m_dListTerms.size(), m_tStream ), m_sError );
		CUDA_TRY ( cudaMemcpyAsync ( m_dListDesc.m_p, m_dListTerms.data(), m_dListTerms.size()*sizeof(DevLeaf_t), cudaMemcpyHostToDevice, m_tStream ), m_sError );
		CUDA_TRY ( m_dListBlkStartDev.AllocAsync ( m_dListBlkStart.size(), m_tStream ), m_sError );
		CUDA_TRY ( cudaMemcpyAsync ( m_dListBlkStartDev.m_p, m_dListBlkStart.data(), m_dListBlkStart.size()*4, cudaMemcpyHostToDevice, m_tStream ), m_sError );
	}
	// ... and with bEagerHot (the one-call paths: prepare + run under one lock) K0 starts right here on the index's second stream,
	// behind everything queued on the batch's stream so far (earlier runs own the store until they end); the rest of this
	// function (work items, the 34 MB of device queries, their upload) then overlaps with the decode instead of preceding it
	if ( bEagerHot && ( !m_dHotTerms.empty() || !m_dListTerms.empty() ) )
	{
		CUDA_TRY ( cudaEventRecord ( m_tEvHotDone, m_tStream ), m_sError );
		CUDA_TRY ( cudaStreamWaitEvent ( pIndex->m_tHotStream, m_tEvHotDone, 0 ), m_sError );
		int iRes = BuildHotStore ( pIndex->m_tHotStream );
		if ( iRes!=MGPU_OK )
			return iRes;
		m_bHotPending = true;
	}
	fnMark ( "store descriptors" );

	// estimated work of a query in its class (decides how many items it is cut into)
	auto fnWork = [&] ( const PlannedQuery_t & p, int c ) -> int64_t
	{
		int64_t iWork = 0;
		if ( c==2 || c==4 )
		{
			for ( int g=0; g<p.m_tDev.m_nGroups; ++g )
			{
				const TermInfo_t * pDrv = p.m_dLeafTerms[p.m_tDev.m_dOps[p.m_tDev.m_dGroupOp0[g]].m_uLeaf];
				iWork += pDrv ? (int64_t)pDrv->m_iDocs*p.m_tDev.m_nLeaves : 0;
			}
			return iWork;
		}
		for ( int l=0; l<p.m_tDev.m_nLeaves; ++l )
			iWork += ( ( c==0 || c==3 || c>=5 ) && p.m_tDev.m_dLeaves[l].m_iHot>=0 ) ? (int64_t)uRows/4 : ( p.m_dLeafTerms[l] ? p.m_dLeafTerms[l]->m_iDocs : 0 );
		return ( c==0 || c==3 || c>=5 ) ? iWork + uRows/16 : iWork;
	};

	for ( int c=0; c<NUM_CLASSES; ++c )
	{
		m_dFirstItem[c] = (int)m_dItems.size();
		if ( dOrder[c].empty() )
			continue;
		int64_t iTotalWork = 0;
		for ( int i : dOrder[c] )
		{
			iTotalWork += fnWork ( m_dPlans[i], c );
			m_dStack[c] = std::max ( m_dStack[c], m_dPlans[i].m_nStack );
		}
		const int iOcc = ( c==2 || c==4 ) ? AndOccupancy ( c==4 ) : ( c==0 || c==3 || c>=5 ) ? StreamOccupancy ( m_dStack[c], c==5 ? m_iOrMode : c==6 ? 2 : 0 ) : EvalOccupancy ( m_dStack[c] );
		const int nMaxCtas = pIndex->m_nSMs*iOcc;
		const int64_t iTarget = std::max<int64_t> ( ( c==0 || c==3 || c>=5 ) ? 262144 : 32768, iTotalWork/( (int64_t)nMaxCtas*4 ) );

		struct Part_t { int m_iQuery; int m_nParts; int64_t m_iCostPerPart; };
		std::vector<Part_t> dParts;
		for ( int i : dOrder[c] )
		{
			const PlannedQuery_t & p = m_dPlans[i];
			const int64_t iWork = fnWork ( p, c );
			int64_t nParts = ( iWork+iTarget-1 )/iTarget;
			int iCap = std::max ( 1, 131072/std::max ( 1, p.m_tDev.m_iMaxMatches ) );
			int64_t nUnits = nTiles;
			if ( c==2 || c==4 )
			{
				nUnits = 0;
				for ( int g=0; g<p.m_tDev.m_nGroups; ++g )
				{
					const TermInfo_t * pDrv = p.m_dLeafTerms[p.m_tDev.m_dOps[p.m_tDev.m_dGroupOp0[g]].m_uLeaf];
					nUnits += pDrv ? ( pDrv->m_nBlocks+AND_CHUNK-1 )/AND_CHUNK : 0;
				}
				nUnits = std::max<int64_t> ( nUnits, p.m_tDev.m_nGroups );
			}
			if ( c>=5 )
			{
				// fixed rowid ranges (OR_RANGE_TILES tiles) for every query of the class: the kernel takes the items range by range, so
				// concurrent CTAs read the same rows of the dense store (L2 hits) and later ranges inherit the query's K-th-best bound
				// (small batches: finer ranges, so that the class still fills the GPU; the items of a query share its K-th-best bound)
				const int64_t iRangeTiles = std::max<int64_t> ( 64, std::min<int64_t> ( OR_RANGE_TILES, (int64_t)nTiles*(int64_t)dOrder[c].size()/( 4*(int64_t)nMaxCtas ) ) );
				nParts = std::max<int64_t> ( 1, std::min<int64_t> ( { ( nTiles+iRangeTiles-1 )/iRangeTiles, (int64_t)iCap, nUnits } ) );
			} else
			nParts = std::max<int64_t> ( ( c==2 || c==4 ) ? p.m_tDev.m_nGroups : 1, std::min<int64_t> ( nParts, std::min<int64_t> ( { nUnits, 64, std::max ( iCap, ( c==2 || c==4 ) ? p.m_tDev.m_nGroups : 1 ) } ) ) );
			dParts.push_back ( { i, (int)nParts, iWork/nParts } );
		}
		std::stable_sort ( dParts.begin(), dParts.end(), [] ( const Part_t & a, const Part_t & b ) { return a.m_iCostPerPart>b.m_iCostPerPart; } );

		for ( const Part_t & t : dParts )
		{
			// (the 3 KB device query itself is copied once, in parallel, straight into the pinned upload buffer below)
			const DevQueryCore_t & q = m_dPlans[t.m_iQuery].m_tDev;
			DevSlot_t tSlot { t.m_iQuery, (int)m_dItems.size(), t.m_nParts };
			const int iDevQuery = (int)m_dSlots.size();
			int64_t nUnits = nTiles, iUnit = TILE_W, iLimit = uRows;
			if ( c==2 || c==4 )
			{
				// items = ranges of each group's driver blocks; the parts are shared out between the groups by their block counts
				int64_t dBlocks[MAX_GROUPS], nTotalBlocks = 0;
				for ( int g=0; g<q.m_nGroups; ++g )
				{
					const TermInfo_t * pDrv = m_dPlans[t.m_iQuery].m_dLeafTerms[q.m_dOps[q.m_dGroupOp0[g]].m_uLeaf];
					dBlocks[g] = pDrv ? pDrv->m_nBlocks : 0;
					nTotalBlocks += dBlocks[g];
				}
				int nLeft = t.m_nParts - q.m_nGroups;	// every group gets one item, the rest go by size
				int nMade = 0;
				for ( int g=0; g<q.m_nGroups; ++g )
				{
					int nMine = 1 + ( nTotalBlocks ? (int)( (int64_t)nLeft*dBlocks[g]/nTotalBlocks ) : 0 );
					const int64_t nChunks = std::max<int64_t> ( 1, ( dBlocks[g]+AND_CHUNK-1 )/AND_CHUNK );
					nMine = (int)std::min<int64_t> ( nMine, nChunks );
					for ( int p=0; p<nMine; ++p )
					{
						DevWorkItem_t tItem;
						tItem.m_uQuery = (uint32_t)iDevQuery;
						tItem.m_uRowLo = (uint32_t)( nChunks*p/nMine*AND_CHUNK );
						tItem.m_uRowHi = (uint32_t)std::min<int64_t> ( nChunks*( p+1 )/nMine*AND_CHUNK, dBlocks[g] );
						tItem.m_uPad = (uint32_t)g;
						m_dItems.push_back ( tItem );
						++nMade;
					}
				}
				tSlot.m_nItems = nMade;
				iMaxKeysPerQuery = std::max ( iMaxKeysPerQuery, nMade*q.m_iMaxMatches );
				m_dSlots.push_back ( tSlot );
				m_dDevToQuery.push_back ( t.m_iQuery );
				continue;
			}
			for ( int p=0; p<t.m_nParts; ++p )
			{
				uint64_t uT0 = (uint64_t)nUnits*p/t.m_nParts, uT1 = (uint64_t)nUnits*( p+1 )/t.m_nParts;
				DevWorkItem_t tItem;
				tItem.m_uQuery = (uint32_t)iDevQuery;
				tItem.m_uRowLo = (uint32_t)( uT0*iUnit );
				tItem.m_uRowHi = (uint32_t)std::min<uint64_t> ( uT1*iUnit, (uint64_t)iLimit );
				tItem.m_uPad = 0;
				m_dItems.push_back ( tItem );
			}
			iMaxKeysPerQuery = std::max ( iMaxKeysPerQuery, t.m_nParts*q.m_iMaxMatches );
			m_dSlots.push_back ( tSlot );
			m_dDevToQuery.push_back ( t.m_iQuery );
		}
		m_dCtas[c] = std::min ( nMaxCtas, (int)m_dItems.size()-m_dFirstItem[c] );
		if ( c>=5 )
		{
			const int iFirst = m_dFirstItem[c], n = (int)m_dItems.size()-iFirst;
			std::vector<int32_t> & dOrd = m_dItemOrder[c-5];
			dOrd.resize ( n );
			for ( int i=0; i<n; ++i )
				dOrd[i] = i;
			std::stable_sort ( dOrd.begin(), dOrd.end(), [&] ( int a, int b ) { return m_dItems[iFirst+a].m_uRowLo<m_dItems[iFirst+b].m_uRowLo; } );
		}
	}
	m_dFirstItem[NUM_CLASSES] = (int)m_dItems.size();
	fnMark ( "classes + items" );

	const int nDevQ = (int)m_dSlots.size();
	m_nDevQueries = nDevQ;
	const int nItems = (int)m_dItems.size();
	cudaStream_t tAllocStream = m_tStream;
	m_iPoolCap = m_iKMax + 32768;	// >= K + what one round of any kernel can push (stream: 8 mini-tiles x 8 warps x 512 rows)
	m_iScratchStride = 2*Pow2Ceil ( iMaxKeysPerQuery );

	CUDA_TRY ( m_dQ.AllocAsync ( nDevQ, tAllocStream ), m_sError );
	CUDA_TRY ( m_dI.AllocAsync ( nItems, tAllocStream ), m_sError );
	CUDA_TRY ( m_dCounter.AllocAsync ( NUM_CLASSES, tAllocStream ), m_sError );
	for ( int i=0; i<2; ++i )
		if ( !m_dItemOrder[i].empty() )
			CUDA_TRY ( m_dOrder[i].AllocAsync ( m_dItemOrder[i].size(), tAllocStream ), m_sError );
	CUDA_TRY ( m_dQueryThr.AllocAsync ( nDevQ, tAllocStream ), m_sError );
	CUDA_TRY ( m_dWork.AllocAsync ( 2, tAllocStream ), m_sError );
	if ( tOpt.m_bStats )
		CUDA_TRY ( m_dDebug.AllocAsync ( 8, tAllocStream ), m_sError );
	m_nPool = (size_t)std::max ( { m_dCtas[0], m_dCtas[1], m_dCtas[2], m_dCtas[3], m_dCtas[4], m_dCtas[5], m_dCtas[6] } )*2*m_iPoolCap;
	m_nHitpos = std::max ( (size_t)m_dCtas[1]*MAX_LEAVES*TILE_W, (size_t)m_dCtas[4]*EVAL_WARPS*MAX_LEAVES*32 );
	m_nPre = (size_t)std::max ( { m_dCtas[0], m_dCtas[1], m_dCtas[3], m_dCtas[5], m_dCtas[6] } )*PRE_BLOCKS*32;
	m_nPreHitpos = (size_t)m_dCtas[1]*PRE_BLOCKS*32;
	CUDA_TRY ( m_dItemKeys.AllocAsync ( (size_t)nItems*m_iKMax, tAllocStream ), m_sError );
	CUDA_TRY ( m_dItemOut.AllocAsync ( nItems, tAllocStream ), m_sError );
	CUDA_TRY ( m_dScratch.AllocAsync ( (size_t)nDevQ*m_iScratchStride, tAllocStream ), m_sError );
	CUDA_TRY ( m_dOutKeys.AllocAsync ( (size_t)nQueries*m_iKMax, tAllocStream ), m_sError );
	CUDA_TRY ( m_dOutDocid.AllocAsync ( (size_t)nQueries*m_iKMax, tAllocStream ), m_sError );
	CUDA_TRY ( m_dOutCount.AllocAsync ( nQueries, tAllocStream ), m_sError );
	CUDA_TRY ( m_dOutTotal.AllocAsync ( nQueries, tAllocStream ), m_sError );
	CUDA_TRY ( m_dOutSlot.AllocAsync ( nDevQ, tAllocStream ), m_sError );
	CUDA_TRY ( cudaMemsetAsync ( m_dOutCount.m_p, 0, (size_t)nQueries*4, tAllocStream ), m_sError );
	CUDA_TRY ( cudaMemsetAsync ( m_dOutTotal.m_p, 0, (size_t)nQueries*8, tAllocStream ), m_sError );

	fnMark ( "device allocations" );
	cudaStream_t s = m_tStream;
	{
		// device queries: gathered by a few threads into the index's pinned upload buffer (grow-only, free again once this call's
		// copy has completed below), then one DMA
		DevQueryCore_t * pStage = (DevQueryCore_t *)pIndex->PinnedUpload ( (size_t)nDevQ*sizeof(DevQueryCore_t) );
		if ( !pStage )
		{
			m_sError = "cudaHostAlloc failed";
			return MGPU_E_NOMEM;
		}
		auto fnFill = [&] ( int iFrom, int iTo )
		{
			for ( int i=iFrom; i<iTo; ++i )
			{
				memcpy ( pStage+i, &m_dPlans[m_dSlots[i].m_iQuery].m_tDev, sizeof(DevQueryCore_t) );
				pStage[i].m_iFirstItem = m_dSlots[i].m_iFirstItem;
				pStage[i].m_nItems = m_dSlots[i].m_nItems;
			}
		};
		int nThreads = std::max ( 1, std::min ( { (int)std::thread::hardware_concurrency(), 8, nDevQ/256 } ) );
		if ( nMaxThreads>0 )
			nThreads = std::min ( nThreads, nMaxThreads );	// (a sharded call runs one of these per shard at the same time)
		if ( nThreads<=1 )
			fnFill ( 0, nDevQ );
		else
		{
			std::vector<std::thread> dThreads;
			for ( int t=0; t<nThreads; ++t )
				dThreads.emplace_back ( fnFill, (int)( (int64_t)nDevQ*t/nThreads ), (int)( (int64_t)nDevQ*( t+1 )/nThreads ) );
			for ( auto & t : dThreads )
				t.join();
		}
		CUDA_TRY ( cudaMemcpyAsync ( m_dQ.m_p, pStage, (size_t)nDevQ*sizeof(DevQueryCore_t), cudaMemcpyHostToDevice, s ), m_sError );
		// the few queries with filters / sort keys / hit-level nodes: their extensions (a sharded call: the template's, shared by the shards)
		const std::vector<DevQueryExt_t> & dExt = pTemplateExt ? *pTemplateExt : m_dExt;
		if ( !dExt.empty() )
		{
			CUDA_TRY ( m_dExtDev.AllocAsync ( dExt.size(), s ), m_sError );
			CUDA_TRY ( cudaMemcpyAsync ( m_dExtDev.m_p, dExt.data(), dExt.size()*sizeof(DevQueryExt_t), cudaMemcpyHostToDevice, s ), m_sError );
		}
		m_tStats.h2d_bytes = (int64_t)dExt.size()*sizeof(DevQueryExt_t);
	}
	CUDA_TRY ( cudaMemcpyAsync ( m_dI.m_p, m_dItems.data(), (size_t)nItems*sizeof(DevWorkItem_t), cudaMemcpyHostToDevice, s ), m_sError );
	CUDA_TRY ( cudaMemcpyAsync ( m_dOutSlot.m_p, m_dDevToQuery.data(), (size_t)nDevQ*4, cudaMemcpyHostToDevice, s ), m_sError );
	for ( int i=0; i<2; ++i )
		if ( !m_dItemOrder[i].empty() )
			CUDA_TRY ( cudaMemcpyAsync ( m_dOrder[i].m_p, m_dItemOrder[i].data(), m_dItemOrder[i].size()*4, cudaMemcpyHostToDevice, s ), m_sError );
	CUDA_TRY ( cudaStreamSynchronize ( s ), m_sError );
	fnMark ( "upload + sync" );
	m_tStats.h2d_bytes += (int64_t)nDevQ*sizeof(DevQueryCore_t) + (int64_t)nItems*sizeof(DevWorkItem_t);
	m_tStats.work_items = nItems;

	for ( int c=0; c<NUM_CLASSES; ++c )
	{
		m_tStats.class_queries[c] = (int32_t)dOrder[c].size();
		m_tStats.class_bytes[c] = 0;
		for ( int i : dOrder[c] )
			m_tStats.class_bytes[c] += m_dPlans[i].m_iAlgBytes;
	}
	fnMark ( "events + stats" );
	m_tStats.host_setup_ms = std::chrono::duration<float,std::milli> ( std::chrono::steady_clock::now()-tPlanned ).count();
	return MGPU_OK;
}

/// K0 (+ K0b): decodes the batch's hot keywords once into the dense store and the non-hot keywords of launch class 5 into plain
/// posting lists, on tStream. The store lives in the index's run scratch: contents never survive a run.
int Batch_c::BuildHotStore ( cudaStream_t s )
{
	Index_c * pIndex = m_pIndex;
	Index_c::RunScratch_t & tScr = pIndex->m_tScratch;
	m_tHot = DevHotStore_t {};
	m_tLists = DevPostingLists_t {};
	m_nHotLaunches = 0;
	if ( !m_dHotTerms.empty() )
	{
		CUDA_TRY ( tScr.m_dHotData.Grow ( m_dHotTerms.size()*(size_t)m_iHotStride ), m_sError );
		CUDA_TRY ( tScr.m_dHotEscape.Grow ( (size_t)m_iHotEscapeCap*4+4 ), m_sError );	// { hot slot, rowid, hits, next in bucket }
		CUDA_TRY ( tScr.m_dHotEscapeCount.Grow ( 1+65536 ), m_sError );					// entry counter + the bucket heads (HOT_ESCAPE_BUCKETS)
		if ( m_nHotBitFields )
			CUDA_TRY ( tScr.m_dHotBits.Grow ( m_dHotTerms.size()*(size_t)m_nHotBitFields*(size_t)( m_iHotStride/32 ) ), m_sError );
		if ( m_nListEntries )
		{
			CUDA_TRY ( tScr.m_dListRows.Grow ( m_nListEntries+64 ), m_sError );
			CUDA_TRY ( tScr.m_dListVals.Grow ( m_nListEntries+64 ), m_sError );
		}
		if ( m_nHotLvl )
			CUDA_TRY ( tScr.m_dHotLvlBits.Grow ( (size_t)m_nHotLvl*2*(size_t)( m_iHotStride/32 ) ), m_sError );
	}
	CUDA_TRY ( cudaEventRecord ( m_tEvHot, s ), m_sError );
	DevHotStore_t & tHot = m_tHot;
	if ( !m_dHotTerms.empty() )
	{
		// (the u16 rows are cleared by hot_decode_kernel itself, block by block; only the bitmaps are memset)
		CUDA_TRY ( cudaMemsetAsync ( tScr.m_dHotEscapeCount.m_p, 0, sizeof(int32_t), s ), m_sError );
		CUDA_TRY ( cudaMemsetAsync ( tScr.m_dHotEscapeCount.m_p+1, 0xFF, 65536*sizeof(int32_t), s ), m_sError );	// empty chains
		HotDecodeParams_t H {};
		H.m_tIndex = pIndex->m_tDev;
		H.m_pTerms = m_dHotDesc.m_p;
		H.m_nHot = (int)m_dHotTerms.size();
		H.m_iEscapeCap = m_iHotEscapeCap;
		H.m_pData = tScr.m_dHotData.m_p;
		H.m_pEscape = tScr.m_dHotEscape.m_p;
		H.m_pEscapeCount = tScr.m_dHotEscapeCount.m_p;
		H.m_iStride = m_iHotStride;
		H.m_bTfClass = pIndex->m_tHdr.m_dFields.size()<=4 ? 1 : 0;
		H.m_pBlkStart = m_dHotBlkStartDev.m_p;
		H.m_nBitFields = m_nHotBitFields;
		H.m_pBits = m_nHotBitFields ? tScr.m_dHotBits.m_p : nullptr;
		H.m_pLvlSlot = m_nHotLvl ? m_dHotLvlSlotDev.m_p : nullptr;
		H.m_pLvlBits = m_nHotLvl ? tScr.m_dHotLvlBits.m_p : nullptr;
		if ( m_nHotLvl )
			CUDA_TRY ( cudaMemsetAsync ( tScr.m_dHotLvlBits.m_p, 0, (size_t)m_nHotLvl*2*(size_t)( m_iHotStride/32 )*4, s ), m_sError );
		if ( m_nHotBitFields )
			CUDA_TRY ( cudaMemsetAsync ( tScr.m_dHotBits.m_p, 0, m_dHotTerms.size()*(size_t)m_nHotBitFields*(size_t)( m_iHotStride/32 )*4, s ), m_sError );
		CUDA_TRY ( LaunchHotDecode ( H, pIndex->m_nSMs*8, s ), m_sError );
		++m_nHotLaunches;
		tHot.m_pData = tScr.m_dHotData.m_p;
		tHot.m_pEscape = tScr.m_dHotEscape.m_p;
		tHot.m_pEscapeCount = tScr.m_dHotEscapeCount.m_p;
		tHot.m_iStride = m_iHotStride;
		tHot.m_nHot = (int)m_dHotTerms.size();
		tHot.m_bTfClass = H.m_bTfClass;
		tHot.m_pBits = H.m_pBits;
		tHot.m_iBitStride = m_iHotStride/32;
		tHot.m_nBitFields = m_nHotBitFields;
		tHot.m_pLvlSlot = H.m_pLvlSlot;
		tHot.m_pLvlBits = H.m_pLvlBits;
	}
	if ( !m_dListTerms.empty() )
	{
		// K0b: posting lists of class 5's keywords outside the hot store
		SparseDecodeParams_t L {};
		L.m_tIndex = pIndex->m_tDev;
		L.m_pTerms = m_dListDesc.m_p;
		L.m_pBlkStart = m_dListBlkStartDev.m_p;
		L.m_nTerms = (int)m_dListTerms.size();
		L.m_pRows = tScr.m_dListRows.m_p;
		L.m_pVals = tScr.m_dListVals.m_p;
		CUDA_TRY ( LaunchSparseDecode ( L, pIndex->m_nSMs*8, s ), m_sError );
		++m_nHotLaunches;
		m_tLists.m_pRows = L.m_pRows;
		m_tLists.m_pVals = L.m_pVals;
	}
	CUDA_TRY ( cudaEventRecord ( m_tEvHotDone, s ), m_sError );
	return MGPU_OK;
}

int Batch_c::Run()
{
	if ( !m_nDevQueries )
		return MGPU_OK;
	Index_c * pIndex = m_pIndex;
	CUDA_TRY ( cudaSetDevice ( pIndex->m_iDevice ), m_sError );
	cudaStream_t s = m_tStream;

	CUDA_TRY ( cudaMemsetAsync ( m_dCounter.m_p, 0, NUM_CLASSES*sizeof(int32_t), s ), m_sError );
	CUDA_TRY ( cudaMemsetAsync ( m_dQueryThr.m_p, 0, (size_t)m_nDevQueries*sizeof(unsigned long long), s ), m_sError );
	CUDA_TRY ( cudaMemsetAsync ( m_dWork.m_p, 0, 2*sizeof(unsigned long long), s ), m_sError );
	if ( m_dDebug.m_p )
		CUDA_TRY ( cudaMemsetAsync ( m_dDebug.m_p, 0, 8*sizeof(unsigned long long), s ), m_sError );

	// run-time scratch comes from the index (grow-only, shared by all batches; runs are serialised on the index stream)
	Index_c::RunScratch_t & tScr = pIndex->m_tScratch;
	CUDA_TRY ( tScr.m_dPool.Grow ( m_nPool ), m_sError );
	CUDA_TRY ( tScr.m_dHitpos.Grow ( m_nHitpos ), m_sError );
	CUDA_TRY ( tScr.m_dLeafTf.Grow ( (size_t)m_dCtas[1]*MAX_LEAVES*TILE_W ), m_sError );
	CUDA_TRY ( tScr.m_dPre.Grow ( m_nPre ), m_sError );
	if ( m_dCtas[5] || m_dCtas[6] )
		CUDA_TRY ( tScr.m_dOrList.Grow ( std::max ( (size_t)m_dCtas[5]*StreamOrListCap ( m_iOrMode ), (size_t)m_dCtas[6]*StreamOrListCap ( 2 ) )*EVAL_WARPS ), m_sError );
	CUDA_TRY ( tScr.m_dPreHitpos.Grow ( m_nPreHitpos ), m_sError );

	// K0 / K0b: the batch's hot-term store and posting lists. Prepare ( bEagerHot ) has already started them on the index's second
	// stream (first run only): this stream then just waits for them
	if ( m_bHotPending )
	{
		CUDA_TRY ( cudaStreamWaitEvent ( s, m_tEvHotDone, 0 ), m_sError );
		m_bHotPending = false;
	} else
	{
		int iRes = BuildHotStore ( s );
		if ( iRes!=MGPU_OK )
			return iRes;
	}
	const DevHotStore_t & tHot = m_tHot;
	const DevPostingLists_t & tLists = m_tLists;
	int nLaunches = 1 + m_nHotLaunches;
	CUDA_TRY ( cudaEventRecord ( m_tEv0, s ), m_sError );
	for ( int c=0; c<NUM_CLASSES; ++c )
	{
		const int iFirst = m_dFirstItem[c];
		const int nClassItems = m_dFirstItem[c+1] - iFirst;
		if ( nClassItems<=0 )
			continue;
		EvalParams_t P {};
		P.m_tIndex = pIndex->m_tDev;
		P.m_pQueries = m_dQ.m_p;
		P.m_pQueryExt = m_dExtDev.m_p;
		P.m_pItems = m_dI.m_p + iFirst;
		P.m_nItems = nClassItems;
		P.m_iPoolCap = m_iPoolCap;
		P.m_pPool = tScr.m_dPool.m_p;
		P.m_pItemKeys = m_dItemKeys.m_p + (size_t)iFirst*m_iKMax;
		P.m_pItemOut = m_dItemOut.m_p + iFirst;
		P.m_pCounter = m_dCounter.m_p + c;
		P.m_iKMax = m_iKMax;
		P.m_pHitpos = tScr.m_dHitpos.m_p;
		P.m_pLeafTf = tScr.m_dLeafTf.m_p;
		P.m_pQueryThr = m_dQueryThr.m_p;
		P.m_pOrList = c>=5 ? tScr.m_dOrList.m_p : nullptr;
		P.m_pDebug = c==5 ? m_dDebug.m_p : nullptr;
		P.m_pWork = m_dWork.m_p;
		P.m_tLists = tLists;
		P.m_pItemOrder = ( c>=5 && !m_dItemOrder[c-5].empty() ) ? m_dOrder[c-5].m_p : nullptr;
		P.m_pPre = tScr.m_dPre.m_p;
		P.m_pPreHitpos = tScr.m_dPreHitpos.m_p;
		P.m_tHot = tHot;
		if ( c==2 || c==4 )
			CUDA_TRY ( LaunchAnd ( P, c==4, m_dCtas[c], s ), m_sError );
		else if ( c==0 || c==3 || c>=5 )
			CUDA_TRY ( LaunchStream ( P, m_dStack[c], c==5 ? m_iOrMode : c==6 ? 2 : 0, m_dCtas[c], s ), m_sError );
		else
			CUDA_TRY ( LaunchEval ( P, m_dStack[c], m_dCtas[c], s ), m_sError );
		CUDA_TRY ( cudaEventRecord ( m_dEvClass[c], s ), m_sError );
		m_dClassRan[c] = true;
		++nLaunches;
	}
	CUDA_TRY ( cudaEventRecord ( m_tEv1, s ), m_sError );

	MergeParams_t M {};
	M.m_tIndex = pIndex->m_tDev;
	M.m_pQueries = m_dQ.m_p;
	M.m_pQueryExt = m_dExtDev.m_p;
	M.m_nQueries = m_nDevQueries;
	M.m_iKMax = m_iKMax;
	M.m_pItemKeys = m_dItemKeys.m_p;
	M.m_pItemOut = m_dItemOut.m_p;
	M.m_pScratch = m_dScratch.m_p;
	M.m_iScratchStride = m_iScratchStride;
	M.m_pOutSlot = m_dOutSlot.m_p;
	M.m_pOutKeys = m_dOutKeys.m_p;
	M.m_pOutDocid = m_dOutDocid.m_p;
	M.m_pOutCount = m_dOutCount.m_p;
	M.m_pOutTotal = m_dOutTotal.m_p;
	CUDA_TRY ( LaunchMerge ( M, std::min ( M.m_nQueries, pIndex->m_nSMs*8 ), s ), m_sError );
	CUDA_TRY ( cudaEventRecord ( m_tEv2, s ), m_sError );
	m_tStats.kernel_launches = nLaunches;
	m_bRan = true;
	return MGPU_OK;
}

int Batch_c::Sync()
{
	if ( !m_pIndex )
		return MGPU_OK;
	CUDA_TRY ( cudaSetDevice ( m_pIndex->m_iDevice ), m_sError );
	CUDA_TRY ( cudaStreamSynchronize ( m_tStream ), m_sError );
	if ( m_bRan && m_nDevQueries )
	{
		cudaEventElapsedTime ( &m_tStats.eval_kernel_ms, m_tEv0, m_tEv1 );
		cudaEventElapsedTime ( &m_tStats.hot_decode_ms, m_tEvHot, m_tEvHotDone );
		m_tStats.hot_terms = (int32_t)m_dHotTerms.size();
		cudaEvent_t tPrev = m_tEv0;
		for ( int c=0; c<NUM_CLASSES; ++c )
		{
			m_tStats.class_ms[c] = 0.0f;
			if ( m_dClassRan[c] )
			{
				cudaEventElapsedTime ( &m_tStats.class_ms[c], tPrev, m_dEvClass[c] );
				tPrev = m_dEvClass[c];
			}
		}
		cudaEventElapsedTime ( &m_tStats.merge_kernel_ms, m_tEv1, m_tEv2 );
		{
			unsigned long long dWork[2] = { 0, 0 };
			if ( cudaMemcpy ( dWork, m_dWork.m_p, sizeof(dWork), cudaMemcpyDeviceToHost )==cudaSuccess )
			{
				m_tStats.hitlist_bytes = (int64_t)dWork[0];
				m_tStats.attr_rows = (int64_t)dWork[1];
			}
		}
		m_tStats.or_kernel = m_iOrMode;
		if ( m_dDebug.m_p )
		{
			unsigned long long dDbg[8];
			if ( cudaMemcpy ( dDbg, m_dDebug.m_p, sizeof(dDbg), cudaMemcpyDeviceToHost )==cudaSuccess )
				fprintf ( stderr, "[mgpu stats] class 5: %d queries, mini-tiles %llu, candidate rows hot-only %llu, with sparse postings %llu, present rows %llu, %.2f ms\n",
					m_tStats.class_queries[5], dDbg[0], dDbg[1], dDbg[2], dDbg[4], m_tStats.class_ms[5] );
		}
	}
	return MGPU_OK;
}

int Batch_c::Fetch ( mgpu_result * pResults )
{
	struct FetchTimer_t
	{
		mgpu_batch_stats & m_tStats;
		std::chrono::steady_clock::time_point m_tStart = std::chrono::steady_clock::now();
		~FetchTimer_t() { m_tStats.host_fetch_ms = std::chrono::duration<float,std::milli> ( std::chrono::steady_clock::now()-m_tStart ).count(); }
	} tTimer { m_tStats };
	const int nQueries = (int)m_dPlans.size();
	for ( int i=0; i<nQueries; ++i )
	{
		mgpu_result & r = pResults[i];
		r.status = m_dPlans[i].m_iStatus;
		r.n_matches = 0;
		r.total_found = 0;
		if ( r.word_stats )
			for ( size_t w=0; w<m_dPlans[i].m_dWordStats.size(); ++w )
				r.word_stats[w] = m_dPlans[i].m_dWordStats[w];
	}
	if ( !m_nDevQueries )
		return MGPU_OK;
	const auto tWait = std::chrono::steady_clock::now();
	int iRes = Sync();
	m_tStats.host_wait_ms = std::chrono::duration<float,std::milli> ( std::chrono::steady_clock::now()-tWait ).count();
	if ( iRes!=MGPU_OK )
		return iRes;

	// download through the index's pinned staging buffer (serialised by the index lock)
	std::lock_guard<std::mutex> tGuard ( m_pIndex->m_tLock );
	const size_t nSlots = (size_t)nQueries*m_iKMax;
	const size_t iOffDocid = nSlots*sizeof(Key128_t), iOffTotal = iOffDocid + nSlots*8, iOffCount = iOffTotal + (size_t)nQueries*8;
	uint8_t * pStage = (uint8_t *)m_pIndex->Pinned ( iOffCount + (size_t)nQueries*4 );
	if ( !pStage )
	{
		m_sError = "cudaHostAlloc failed";
		return MGPU_E_NOMEM;
	}
	const Key128_t * dKeys = (const Key128_t *)pStage;
	const int64_t * dDocid = (const int64_t *)( pStage+iOffDocid );
	const int64_t * dTotal = (const int64_t *)( pStage+iOffTotal );
	const int32_t * dCount = (const int32_t *)( pStage+iOffCount );
	cudaStream_t s = m_tStream;
	CUDA_TRY ( cudaMemcpyAsync ( pStage, m_dOutKeys.m_p, nSlots*sizeof(Key128_t), cudaMemcpyDeviceToHost, s ), m_sError );
	CUDA_TRY ( cudaMemcpyAsync ( pStage+iOffDocid, m_dOutDocid.m_p, nSlots*8, cudaMemcpyDeviceToHost, s ), m_sError );
	CUDA_TRY ( cudaMemcpyAsync ( pStage+iOffTotal, m_dOutTotal.m_p, (size_t)nQueries*8, cudaMemcpyDeviceToHost, s ), m_sError );
	CUDA_TRY ( cudaMemcpyAsync ( pStage+iOffCount, m_dOutCount.m_p, (size_t)nQueries*4, cudaMemcpyDeviceToHost, s ), m_sError );
	CUDA_TRY ( cudaStreamSynchronize ( s ), m_sError );
	m_tStats.d2h_bytes = (int64_t)nQueries*12 + (int64_t)nSlots*24;

	auto fnUnpack = [&] ( int iFrom, int iTo )
	{
	for ( int iQuery=iFrom; iQuery<iTo; ++iQuery )
	{
		const PlannedQuery_t & p = m_dPlans[iQuery];
		mgpu_result & r = pResults[iQuery];
		if ( p.m_iStatus!=MGPU_OK )
			continue;
		r.n_matches = dCount[iQuery];
		r.total_found = dTotal[iQuery];
		for ( int i=0; i<r.n_matches; ++i )
		{
			const Key128_t & k = dKeys[(size_t)iQuery*m_iKMax+i];
			uint32_t uGlobalRow = ~(uint32_t)( k.m_uLo>>32 );
			if ( r.rowid ) r.rowid[i] = uGlobalRow - m_pIndex->m_uRowidBase;
			if ( r.weight ) r.weight[i] = (int32_t)(uint32_t)k.m_uLo;
			if ( r.docid ) r.docid[i] = dDocid[(size_t)iQuery*m_iKMax+i];
			if ( r.sort_attr )
			{
				int64_t v = 0;
				if ( p.m_iFirstIntKeyShift>=0 )
				{
					uint64_t u = k.m_uHi>>p.m_iFirstIntKeyShift;
					if ( p.m_iFirstIntKeyBits==32 )
					{
						u &= 0xffffffffull;
						if ( !p.m_bFirstIntKeyDesc ) u = ~u & 0xffffffffull;
						v = (int64_t)u;
					} else
					{
						if ( !p.m_bFirstIntKeyDesc ) u = ~u;
						v = (int64_t)( u ^ 0x8000000000000000ull );
					}
				}
				r.sort_attr[i] = v;
			}
		}
	}
	};
	// 24 B per result slot into the caller's arrays: a few threads for big batches (3.5 ms single-threaded for 10k x 100)
	const int nThreads = (int)std::max<int64_t> ( 1, std::min<int64_t> ( { (int64_t)std::thread::hardware_concurrency(), 8, (int64_t)( nSlots/65536 ) } ) );
	if ( nThreads<=1 )
		fnUnpack ( 0, nQueries );
	else
	{
		std::vector<std::thread> dThreads;
		for ( int t=0; t<nThreads; ++t )
			dThreads.emplace_back ( fnUnpack, (int)( (int64_t)nQueries*t/nThreads ), (int)( (int64_t)nQueries*( t+1 )/nThreads ) );
		for ( auto & t : dThreads )
			t.join();
	}
	return MGPU_OK;
}

int Batch_c::ExportKeys ( void * pDevKeys, void * pDevCounts, void * pDevTotal, int iK, void * pDevDocids )
{
	// device-to-device repack of the per-query keys [nq][KMax] -> [nq][iK]; outputs are already in the caller's query order
	const int nQueries = (int)m_dPlans.size();
	CUDA_TRY ( cudaSetDevice ( m_pIndex->m_iDevice ), m_sError );
	cudaStream_t s = m_nDevQueries ? m_tStream : m_pIndex->m_tStream;
	if ( !m_nDevQueries )
	{
		CUDA_TRY ( cudaMemsetAsync ( pDevCounts, 0, (size_t)nQueries*4, s ), m_sError );
		CUDA_TRY ( cudaMemsetAsync ( pDevTotal, 0, (size_t)nQueries*8, s ), m_sError );
	} else
	{
		const int iW = std::min ( iK, m_iKMax );
		CUDA_TRY ( cudaMemcpy2DAsync ( pDevKeys, (size_t)iK*sizeof(Key128_t), m_dOutKeys.m_p, (size_t)m_iKMax*sizeof(Key128_t),
			(size_t)iW*sizeof(Key128_t), nQueries, cudaMemcpyDeviceToDevice, s ), m_sError );
		if ( pDevDocids )
			CUDA_TRY ( cudaMemcpy2DAsync ( pDevDocids, (size_t)iK*8, m_dOutDocid.m_p, (size_t)m_iKMax*8, (size_t)iW*8, nQueries, cudaMemcpyDeviceToDevice, s ), m_sError );
		CUDA_TRY ( cudaMemcpyAsync ( pDevCounts, m_dOutCount.m_p, (size_t)nQueries*4, cudaMemcpyDeviceToDevice, s ), m_sError );
		CUDA_TRY ( cudaMemcpyAsync ( pDevTotal, m_dOutTotal.m_p, (size_t)nQueries*8, cudaMemcpyDeviceToDevice, s ), m_sError );
	}
	return MGPU_OK;
}

} // namespace mgpu
