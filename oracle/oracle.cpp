// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under manticoresearch_b200/ may
// include, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker / the CPU baseline.
//
// A from-scratch CPU restatement of Manticore Search 3.6.0's full-text query hot path (the reference
// cannot be built here: no bison/flex/boost), following, and citing, the reference files:
//   VByte ...................... src/fileio.cpp:31-45
//   doclist/hitlist/skiplist ... src/sphinx.cpp:374-388, 391-451, 459-549, 13042-13077
//   dictionary (keywords) ...... src/indexformat.cpp:641-691, 331-344
//   eval tree .................. src/searchnode.cpp (ExtTerm 1876-2026, ExtAnd 2570-2700, ExtMultiAnd 2716-3260,
//                                ExtOr 3465-3551, ExtMaybe 3565-3604, ExtAndNot 3618-3711, ExtNWay 3767-3848,
//                                FSMphrase 3884-3953, FSMproximity 3958-4075, factory 1599-1811)
//   rankers / IDF .............. src/sphinxsearch.cpp:1033-1169, 1197-1437, 4167-4380
//   sorter ..................... src/sphinxsort.cpp:722-761, 4534-4790
//   shard merge ................ src/searchd.cpp:3897-3952, 4653-4738
// Parity pinning: tests/test_oracle_golden.py checks this file against the reference's own golden
// vectors (test/test_019, test_037, test_322, test_116, test_114 model.bin; gtests_rtstuff.cpp:244-335).
//
// Doc-at-a-time pull iterators are used instead of the reference's 32-doc chunks; the asymptotics are
// the same (skiplist-assisted AdvanceTo, rarest-first leapfrog AND, 2-way merges, binary-heap top-K).

#include "../include/mgpu.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace
{

typedef uint8_t BYTE;
typedef uint32_t DWORD;
typedef uint16_t WORD;
typedef uint32_t RowID_t;
typedef uint32_t Hitpos_t;
static const RowID_t INVALID_ROWID = 0xFFFFFFFFu;
static const Hitpos_t EMPTY_HIT = 0;

#define SPH_BM25_K1 1.2f		// src/searchnode.cpp:45
#define SPH_BM25_SCALE 1000		// src/sphinxsearch.cpp:31

static inline int HitField ( Hitpos_t u )				{ return (int)( u>>24 ); }
static inline DWORD HitPosWithField ( Hitpos_t u )		{ return u & ~( 1u<<23 ); }

//////////////////////////////////////////////////////////////////////////
// files
//////////////////////////////////////////////////////////////////////////

struct Mapped_t
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
	~Mapped_t() { if ( m_p ) munmap ( (void*)m_p, m_iLen ); }
};

// src/fileio.cpp:31-45
static inline uint64_t Unzip ( const BYTE * & p )
{
	DWORD b = *p++;
	uint64_t res = 0;
	while ( b & 0x80 )
	{
		res = ( res<<7 ) + ( b & 0x7f );
		b = *p++;
	}
	return ( res<<7 ) + b;
}

struct Cursor_t
{
	const BYTE * m_p;
	DWORD Dword()			{ DWORD v; memcpy ( &v, m_p, 4 ); m_p += 4; return v; }
	int64_t Offset()		{ int64_t v; memcpy ( &v, m_p, 8 ); m_p += 8; return v; }
	BYTE Byte()				{ return *m_p++; }
	std::string String()	{ DWORD n = Dword(); std::string s ( (const char*)m_p, n ); m_p += n; return s; }
	void FileInfo()			{ m_p += 8+8+8+4; }
};

struct Attr_t
{
	std::string m_sName;
	DWORD m_iBitOffset, m_iBitCount;
};

struct WordEntry_t
{
	int64_t m_iDoclistOffset = 0;
	int m_iDocs = 0;
	int m_iHits = 0;
	int64_t m_iSkiplistOffset = 0;
};

struct SkiplistEntry_t		// src/sphinxsearch.h:35-40
{
	RowID_t m_tBaseRowIDPlus1;
	int64_t m_iOffset;
	int64_t m_iBaseHitlistPos;
};

struct Index_t
{
	Mapped_t m_tSpd, m_tSpp, m_tSpe, m_tSpi, m_tSpa, m_tSpm;
	std::vector<std::string> m_dFields;
	std::vector<Attr_t> m_dAttrs;
	int m_iStride = 0;
	int64_t m_iTotalDocs = 0;
	int64_t m_iRows = 0;
	bool m_bInlineHits = true;
	int m_iSkipBlock = 32;
	std::unordered_map<std::string,WordEntry_t> m_hWords;
	bool m_bWordDict = true;
	std::unordered_map<uint64_t,WordEntry_t> m_hWordIds;	// dict=crc: by word id
	std::string m_sError;

	// the query side of a dict=crc index hashes the keyword with sphFNV64 (src/fnv64.cpp:16-50; CSphDictCRC<false>::GetWordID,
	// src/sphinx.cpp:17318, 17569) and looks the id up (CWordlist::GetWord, src/indexformat.cpp:425-473)
	static uint64_t FNV64 ( const char * s )
	{
		uint64_t h = 14695981039346656037ULL;
		while ( *s )
		{
			h ^= (BYTE)*s++;
			h *= 1099511628211ULL;
		}
		return h;
	}
	const WordEntry_t * FindWord ( const char * szWord ) const
	{
		if ( m_bWordDict )
		{
			auto it = m_hWords.find ( szWord );
			return it==m_hWords.end() ? nullptr : &it->second;
		}
		auto it = m_hWordIds.find ( FNV64 ( szWord ) );
		return it==m_hWordIds.end() ? nullptr : &it->second;
	}

	bool Open ( const std::string & sPrefix );
	const DWORD * Row ( RowID_t r ) const	{ return (const DWORD *)m_tSpa.m_p + (int64_t)r*m_iStride; }	// src/sphinx.cpp:11984
	int64_t GetAttr ( RowID_t r, int iAttr ) const
	{
		const Attr_t & a = m_dAttrs[iAttr];
		const DWORD * p = Row ( r ) + a.m_iBitOffset/32;
		if ( a.m_iBitCount==64 )
			return (int64_t)( (uint64_t)p[0] | ( (uint64_t)p[1]<<32 ) );
		if ( a.m_iBitCount==32 )
			return (int64_t)p[0];
		return (int64_t)( ( p[0]>>( a.m_iBitOffset & 31 ) ) & ( ( 1u<<a.m_iBitCount )-1 ) );
	}
	bool IsDead ( RowID_t r ) const
	{
		if ( !m_tSpm.m_p || (size_t)( r>>3 )>=m_tSpm.m_iLen )
			return false;
		return ( ( (const DWORD*)m_tSpm.m_p )[r>>5]>>( r & 31 ) ) & 1;
	}
};

bool Index_t::Open ( const std::string & sPrefix )
{
	Mapped_t tSph;
	if ( !tSph.Map ( sPrefix+".sph" ) || !m_tSpd.Map ( sPrefix+".spd" ) || !m_tSpp.Map ( sPrefix+".spp" )
		|| !m_tSpe.Map ( sPrefix+".spe" ) || !m_tSpi.Map ( sPrefix+".spi" ) || !m_tSpa.Map ( sPrefix+".spa" ) )
	{
		m_sError = "failed to map index files at " + sPrefix;
		return false;
	}
	m_tSpm.Map ( sPrefix+".spm" );

	// LoadHeader, src/sphinx.cpp:13252-13392
	Cursor_t c { tSph.m_p };
	if ( c.Dword()!=0x58485053 ) { m_sError = "bad magic"; return false; }
	DWORD uVer = c.Dword();
	if ( uVer<57 || uVer>62 ) { m_sError = "bad version"; return false; }
	DWORD nFields = c.Dword();
	for ( DWORD i=0; i<nFields; ++i )
	{
		m_dFields.push_back ( c.String() );
		c.Dword(); c.Byte();
	}
	DWORD nAttrs = c.Dword();
	DWORD uMaxBit = 0;
	for ( DWORD i=0; i<nAttrs; ++i )
	{
		Attr_t a;
		a.m_sName = c.String();
		c.Dword(); c.Dword();
		a.m_iBitOffset = c.Dword();
		a.m_iBitCount = c.Dword();
		c.Byte();
		if ( uVer>=61 )
			c.Dword();
		uMaxBit = std::max ( uMaxBit, a.m_iBitOffset+a.m_iBitCount );
		m_dAttrs.push_back ( a );
	}
	m_iStride = (int)( ( uMaxBit+31 )/32 );
	int64_t iCpOffset = c.Offset();
	DWORD nCp = c.Dword();
	c.Byte(); c.Dword(); c.Dword();
	m_iTotalDocs = c.Dword();
	c.Offset();
	// LoadIndexSettings, :13207-13249
	c.Dword(); c.Dword(); c.Dword();
	c.Byte(); c.String(); c.String();
	c.Byte();
	DWORD uHitless = c.Dword();
	DWORD uHitFormat = c.Dword();
	c.Byte(); c.String();
	c.Dword(); c.Dword(); c.Dword(); c.Dword();
	c.Byte(); c.String();
	c.Byte(); c.Byte(); c.String(); c.String(); c.Offset();
	m_iSkipBlock = (int)c.Dword();
	if ( uVer>=60 )
		c.String();
	if ( uHitless!=0 ) { m_sError = "hitless indexes not supported"; return false; }
	m_bInlineHits = ( uHitFormat==1 );
	// tokenizer settings, src/indexsettings.cpp:303-333
	c.Byte(); c.String(); c.Dword();
	if ( c.Byte() ) { DWORD n = c.Dword(); while ( n-- ) c.String(); }
	c.String(); c.FileInfo();
	c.String(); c.String(); c.Dword(); c.String(); c.String(); c.String();
	// dict settings, src/indexsettings.cpp:405-453
	c.String(); c.String();
	if ( c.Byte() ) { DWORD n = c.Dword(); while ( n-- ) Unzip ( c.m_p ); }
	c.String();
	{ DWORD n = c.Dword(); while ( n-- ) { c.String(); c.FileInfo(); } }
	if ( c.Byte() ) { DWORD n = c.Dword(); while ( n-- ) c.String(); }
	{ DWORD n = c.Dword(); while ( n-- ) { c.String(); c.FileInfo(); } }
	c.Dword();
	bool bWordDict = c.Byte()!=0;
	c.Byte(); c.String();
	m_iRows = c.Offset();
	m_bWordDict = bWordDict;
	if ( !bWordDict )
	{
		// dict=crc: checkpoints are { u64 first word id, u64 offset } (CSphDiskDictTraits::DictEnd, src/sphinx.cpp:18269-18286); each points
		// at up to 64 entries of { zipped word id delta, zipped doclist offset delta, docs, hits, [skiplist offset] }, closed by a zero
		// delta + the last doclist's length (DictEntry / DictEndEntries, :18288-18337; read back by CWordlist::GetWord, src/indexformat.cpp:425-473)
		const BYTE * pCp = m_tSpi.m_p + iCpOffset;
		for ( DWORD i=0; i<nCp; ++i, pCp += 16 )
		{
			int64_t iOff; memcpy ( &iOff, pCp+8, 8 );
			const BYTE * p = m_tSpi.m_p + iOff;
			uint64_t uID = 0;
			int64_t iDoclist = 0;
			while ( true )
			{
				uint64_t uDelta = Unzip ( p );
				if ( !uDelta )
					break;
				uID += uDelta;
				iDoclist += (int64_t)Unzip ( p );
				WordEntry_t e;
				e.m_iDoclistOffset = iDoclist;
				e.m_iDocs = (int)Unzip ( p );
				e.m_iHits = (int)Unzip ( p );
				if ( e.m_iDocs>m_iSkipBlock )
					e.m_iSkiplistOffset = (int64_t)Unzip ( p );
				m_hWordIds.emplace ( uID, e );
			}
		}
		return true;
	}

	// dictionary: checkpoints then 64-word blocks, src/indexformat.cpp:331-344, 641-691
	const BYTE * pCp = m_tSpi.m_p + iCpOffset;
	for ( DWORD i=0; i<nCp; ++i )
	{
		DWORD n; memcpy ( &n, pCp, 4 ); pCp += 4+n;
		int64_t iOff; memcpy ( &iOff, pCp, 8 ); pCp += 8;
		const BYTE * p = m_tSpi.m_p + iOff;
		char sWord[1024];
		int iLen = 0;
		while ( true )
		{
			BYTE uPack = *p++;
			if ( !uPack )
				break;
			int iMatch, iDelta;
			if ( uPack & 0x80 ) { iDelta = ( ( uPack>>4 ) & 7 )+1; iMatch = uPack & 15; }
			else { iDelta = uPack & 127; iMatch = *p++; }
			memcpy ( sWord+iMatch, p, iDelta );
			p += iDelta;
			iLen = iMatch+iDelta;
			WordEntry_t e;
			e.m_iDoclistOffset = (int64_t)Unzip ( p );
			e.m_iDocs = (int)Unzip ( p );
			e.m_iHits = (int)Unzip ( p );
			if ( e.m_iDocs>=256 )
				p++;	// doclist size hint
			if ( e.m_iDocs>m_iSkipBlock )
				e.m_iSkiplistOffset = (int64_t)Unzip ( p );
			m_hWords.emplace ( std::string ( sWord, iLen ), e );
		}
	}
	return true;
}

//////////////////////////////////////////////////////////////////////////
// DiskIndexQword_c, src/sphinx.cpp:357-550
//////////////////////////////////////////////////////////////////////////

struct Qword_t
{
	const Index_t * m_pIndex = nullptr;
	std::string m_sWord;
	int m_iDocs = 0, m_iHits = 0;
	int64_t m_iOrderDocs = -1;			// what orders the keywords of multi-keyword nodes (GetDocsCount): the dictionary's count (the reference),
										// or the whole index's count when this index is one rowid-range shard of it (mgpu_query::shard_of_global)
	int64_t OrderDocs() const			{ return m_iOrderDocs>=0 ? m_iOrderDocs : m_iDocs; }
	int m_iAtomPos = 0;
	float m_fBoost = 1.0f;
	bool m_bExcluded = false, m_bExpanded = false;

	std::vector<SkiplistEntry_t> m_dSkiplist;
	int m_iSkipListBlock = -1;
	const BYTE * m_pDoc = nullptr;		// doclist read cursor
	int64_t m_iDoclistOffset = 0;

	RowID_t m_tRowID = INVALID_ROWID;
	DWORD m_uMatchHits = 0;
	DWORD m_uFields = 0;				// m_dQwordFields.GetMask32()
	uint64_t m_iHitlistPos = 0;
	uint64_t m_uHitPosition = 0;

	// hit decoder
	const BYTE * m_pHit = nullptr;
	int m_uHitState = 0;
	DWORD m_uInlinedHit = 0;
	Hitpos_t m_iHitPos = 0;

	// DiskIndexQwordSetup_c::Setup, :12950-13078
	bool Setup ( const Index_t * pIndex, const std::string & sWord )
	{
		m_pIndex = pIndex;
		m_sWord = sWord;
		const WordEntry_t * pEntry = pIndex->FindWord ( sWord.c_str() );
		if ( !pEntry )
			return false;
		const WordEntry_t & e = *pEntry;
		m_iDocs = e.m_iDocs;
		m_iHits = e.m_iHits;
		m_iDoclistOffset = e.m_iDoclistOffset;
		const int iBlk = pIndex->m_iSkipBlock;
		if ( e.m_iDocs>iBlk )
		{
			const BYTE * pSkip = pIndex->m_tSpe.m_p + e.m_iSkiplistOffset;
			m_dSkiplist.push_back ( { 0, e.m_iDoclistOffset, 0 } );
			for ( int i=1; i<( m_iDocs/iBlk ); ++i )
			{
				SkiplistEntry_t p = m_dSkiplist.back();
				SkiplistEntry_t t;
				t.m_tBaseRowIDPlus1 = p.m_tBaseRowIDPlus1 + iBlk + (DWORD)Unzip ( pSkip );
				t.m_iOffset = p.m_iOffset + 4*iBlk + (int64_t)Unzip ( pSkip );
				t.m_iBaseHitlistPos = p.m_iBaseHitlistPos + (int64_t)Unzip ( pSkip );
				m_dSkiplist.push_back ( t );
			}
		}
		m_pDoc = pIndex->m_tSpd.m_p + e.m_iDoclistOffset;
		return true;
	}

	// ReadNext, :511-549
	void ReadNext()
	{
		DWORD uDelta = (DWORD)Unzip ( m_pDoc );
		if ( !uDelta )
		{
			m_tRowID = INVALID_ROWID;
			m_pDoc--;	// stay on the terminator
			return;
		}
		m_tRowID += uDelta;
		if ( m_pIndex->m_bInlineHits )
		{
			m_uMatchHits = (DWORD)Unzip ( m_pDoc );
			const DWORD uFirst = (DWORD)Unzip ( m_pDoc );
			if ( m_uMatchHits==1 )
			{
				DWORD uField = (DWORD)Unzip ( m_pDoc );
				m_iHitlistPos = (uint64_t)uFirst | ( (uint64_t)uField<<23 ) | ( 1ull<<63 );
				DWORD iField = ( uField>>1 ) & 255;
				m_uFields = iField<32 ? ( 1u<<iField ) : 0;
			} else
			{
				m_uFields = uFirst;
				m_uHitPosition += Unzip ( m_pDoc );
				m_iHitlistPos = m_uHitPosition;
			}
		} else
		{
			m_iHitlistPos += Unzip ( m_pDoc );
			m_uFields = (DWORD)Unzip ( m_pDoc );
			m_uMatchHits = (DWORD)Unzip ( m_pDoc );
		}
	}

	// FindSpan (sphinxstd.h): last entry with base <= rowid
	static int FindSpan ( const SkiplistEntry_t * p, int n, RowID_t tRef )
	{
		if ( !n || tRef<p[0].m_tBaseRowIDPlus1 )
			return -1;
		if ( tRef>=p[n-1].m_tBaseRowIDPlus1 )
			return n-1;
		int l = 0, r = n-1;
		while ( r-l>1 )
		{
			int m = ( l+r )/2;
			if ( tRef<p[m].m_tBaseRowIDPlus1 ) r = m; else l = m;
		}
		return l;
	}

	// HintRowID, :407-451
	bool HintRowID ( RowID_t tRowID )
	{
		const int n = (int)m_dSkiplist.size();
		if ( m_iSkipListBlock==-1 )
		{
			m_iSkipListBlock = FindSpan ( m_dSkiplist.data(), n, tRowID );
			if ( m_iSkipListBlock<0 )
				return false;
		} else
		{
			if ( m_iSkipListBlock<n-1 )
			{
				int iNext = m_iSkipListBlock+1;
				if ( tRowID>=m_dSkiplist[iNext].m_tBaseRowIDPlus1 )
				{
					int iRes = FindSpan ( &m_dSkiplist[iNext], n-iNext, tRowID );
					if ( iRes<0 )
						return false;
					m_iSkipListBlock = iRes+iNext;
				}
			} else
				return false;
		}
		const SkiplistEntry_t & t = m_dSkiplist[m_iSkipListBlock];
		if ( t.m_iOffset<=(int64_t)( m_pDoc-m_pIndex->m_tSpd.m_p ) )
			return false;
		m_pDoc = m_pIndex->m_tSpd.m_p + t.m_iOffset;
		m_tRowID = t.m_tBaseRowIDPlus1-1;
		m_uHitPosition = m_iHitlistPos = (uint64_t)t.m_iBaseHitlistPos;
		return true;
	}

	// AdvanceTo, :391-404
	RowID_t AdvanceTo ( RowID_t tRowID )
	{
		if ( m_tRowID!=INVALID_ROWID && tRowID<=m_tRowID )
			return m_tRowID;
		bool bRewound = HintRowID ( tRowID );
		if ( bRewound || m_tRowID==INVALID_ROWID )
			ReadNext();
		while ( m_tRowID<tRowID )
			ReadNext();
		return m_tRowID;
	}

	// SeekHitlist / GetNextHit, :459-501, 374-388
	void SeekHitlist ( uint64_t uOff )
	{
		if ( uOff>>63 )
		{
			m_uHitState = 1;
			m_uInlinedHit = (DWORD)uOff;
		} else
		{
			m_uHitState = 0;
			m_iHitPos = EMPTY_HIT;
			m_pHit = m_pIndex->m_tSpp.m_p + uOff;
		}
	}
	Hitpos_t GetNextHit()
	{
		switch ( m_uHitState )
		{
		case 0:
			{
				DWORD iDelta = (DWORD)Unzip ( m_pHit );
				if ( iDelta ) m_iHitPos += iDelta; else m_iHitPos = EMPTY_HIT;
				return m_iHitPos;
			}
		case 1:		m_uHitState = 2; return m_uInlinedHit;
		default:	m_uHitState = 0; return EMPTY_HIT;
		}
	}
};

//////////////////////////////////////////////////////////////////////////
// eval tree
//////////////////////////////////////////////////////////////////////////

struct ExtDoc_t { RowID_t m_tRowID; DWORD m_uDocFields; float m_fTFIDF; };	// src/sphinxint.h:738-743
struct ExtHit_t		// src/sphinxint.h:725-736
{
	RowID_t m_tRowID; Hitpos_t m_uHitpos; WORD m_uQuerypos; WORD m_uNodepos; WORD m_uSpanlen; WORD m_uMatchlen; DWORD m_uWeight; DWORD m_uQposMask;
};

struct ExtQword_t	// src/searchnode.h:34-45
{
	std::string m_sWord;
	int m_iDocs, m_iHits, m_iQueryPos;
	float m_fIDF, m_fBoost;
	bool m_bExpanded, m_bExcluded;
	int m_iFirstWordIdx;	// index in mgpu_query.words[] of the first occurrence (for stats output)
};
typedef std::unordered_map<std::string,ExtQword_t> QwordsHash_t;

// sphSort for <=33 elements is this insertion sort (src/sphinxstd.h:853-866): NOT stable, equal keys end reversed
template<typename T, typename LESS>
static void RefSort ( std::vector<T> & d, LESS fnLess )
{
	if ( d.size()>33 )
	{
		std::stable_sort ( d.begin(), d.end(), fnLess );	// never reached by the golden/parity query sets; see DESIGN.md
		return;
	}
	for ( size_t i=1; i<d.size(); ++i )
		for ( size_t j=i; j>0; --j )
		{
			if ( fnLess ( d[j-1], d[j] ) )
				break;
			std::swap ( d[j], d[j-1] );
		}
}

struct Node_c
{
	int m_iAtomPos = 0;
	bool m_bQPosReverse = false;
	virtual ~Node_c() {}
	/// next document in ascending rowid order; false at the end
	virtual bool Next ( ExtDoc_t & tDoc ) = 0;
	/// non-binding skip hint (ExtNode_i::HintRowID)
	virtual void HintRowID ( RowID_t ) {}
	/// appends the hits of the document most recently returned by Next(), ordered by (hitpos, qpos)
	virtual void CollectHits ( std::vector<ExtHit_t> & dHits ) = 0;
	virtual int GetDocsCount() { return INT_MAX; }
	virtual int GetQwords ( QwordsHash_t & hQwords ) = 0;
	virtual void SetQwordsIDF ( const QwordsHash_t & hQwords ) = 0;
};

static int RegisterQword ( Qword_t & q, bool bNotWeighted, float & fIDF, QwordsHash_t & hQwords, int iWordIdx )
{
	// ExtTerm_T::GetQwords, src/searchnode.cpp:2030-2057
	fIDF = 0.0f;
	auto it = hQwords.find ( q.m_sWord );
	if ( !bNotWeighted && it!=hQwords.end() && !it->second.m_bExcluded )
		it->second.m_iQueryPos = std::min ( it->second.m_iQueryPos, q.m_iAtomPos );
	if ( bNotWeighted || it!=hQwords.end() )
		return q.m_bExcluded ? -1 : q.m_iAtomPos;
	fIDF = -1.0f;
	ExtQword_t t;
	t.m_sWord = q.m_sWord;
	t.m_iDocs = q.m_iDocs;
	t.m_iHits = q.m_iHits;
	t.m_iQueryPos = q.m_iAtomPos;
	t.m_fIDF = -1.0f;
	t.m_fBoost = q.m_fBoost;
	t.m_bExpanded = q.m_bExpanded;
	t.m_bExcluded = q.m_bExcluded;
	t.m_iFirstWordIdx = iWordIdx;
	hQwords.emplace ( q.m_sWord, t );
	return q.m_bExcluded ? -1 : q.m_iAtomPos;
}

/// ExtTerm_T, src/searchnode.cpp:1876-2026
struct TermNode_c : Node_c
{
	Qword_t m_tQword;
	DWORD m_uQueriedFields = 0xFFFFFFFFu;
	bool m_bNotWeighted = false;
	bool m_bUseBM25 = true;
	float m_fIDF = 0.0f;
	int m_iWordIdx = 0;
	uint64_t m_uCurHitlistPos = 0;
	RowID_t m_tCurRowID = INVALID_ROWID;
	int m_iTermPos = 0;			// TermPosFilter_e: 0 none, 1 field start, 2 field end, 3 both, 4 field limit (src/searchnode.cpp:875-878, 1145-1146)
	int m_iMaxFieldPos = 0;

	// TermAcceptor_T<..>::IsAcceptableHit, src/searchnode.cpp:2264-2285
	bool IsAcceptableHit ( Hitpos_t uHit ) const
	{
		const DWORD uPos = uHit & 0x7FFFFFu;
		const bool bEnd = ( uHit>>23 ) & 1u;
		switch ( m_iTermPos )
		{
		case 1:		return uPos==1;
		case 2:		return bEnd;
		case 3:		return uPos==1 && bEnd;
		case 4:		return (int)uPos<=m_iMaxFieldPos;
		default:	return true;
		}
	}
	bool HitFits ( Hitpos_t uHit ) const
	{
		int iField = HitField ( uHit );
		return iField<32 && ( m_uQueriedFields & ( 1u<<iField ) ) && IsAcceptableHit ( uHit );
	}

	bool Next ( ExtDoc_t & tDoc ) override
	{
		if ( !m_tQword.m_iDocs )
			return false;
		while ( true )
		{
			m_tQword.ReadNext();
			if ( m_tQword.m_tRowID==INVALID_ROWID )
			{
				m_tQword.m_iDocs = 0;
				return false;
			}
			if ( !( m_tQword.m_uFields & m_uQueriedFields ) )
				continue;
			if ( m_iTermPos )
			{
				// ExtConditional_T::GetDocsChunk, src/searchnode.cpp:2331-2400: the document survives iff it has an acceptable hit
				bool bAny = false;
				m_tQword.SeekHitlist ( m_tQword.m_iHitlistPos );
				for ( Hitpos_t uHit = m_tQword.GetNextHit(); uHit!=EMPTY_HIT; uHit = m_tQword.GetNextHit() )
					bAny |= HitFits ( uHit );
				if ( !bAny )
					continue;
			}
			tDoc.m_tRowID = m_tQword.m_tRowID;
			tDoc.m_uDocFields = m_tQword.m_uFields & m_uQueriedFields;
			tDoc.m_fTFIDF = 0.0f;
			if ( m_bUseBM25 )
				tDoc.m_fTFIDF = float(m_tQword.m_uMatchHits) / float(m_tQword.m_uMatchHits+SPH_BM25_K1) * m_fIDF;	// :1946
			m_uCurHitlistPos = m_tQword.m_iHitlistPos;
			m_tCurRowID = tDoc.m_tRowID;
			return true;
		}
	}
	void HintRowID ( RowID_t tRowID ) override	{ if ( m_tQword.m_iDocs ) m_tQword.HintRowID ( tRowID ); }
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override
	{
		m_tQword.SeekHitlist ( m_uCurHitlistPos );
		while ( true )
		{
			Hitpos_t uHit = m_tQword.GetNextHit();
			if ( uHit==EMPTY_HIT )
				break;
			if ( !HitFits ( uHit ) )
				continue;
			dHits.push_back ( { m_tCurRowID, uHit, (WORD)m_iAtomPos, 0, 1, 1, 1, 0 } );
		}
	}
	int GetDocsCount() override			{ return (int)std::min<int64_t> ( m_tQword.OrderDocs(), INT_MAX-1 ); }
	int GetQwords ( QwordsHash_t & h ) override			{ return RegisterQword ( m_tQword, m_bNotWeighted, m_fIDF, h, m_iWordIdx ); }
	void SetQwordsIDF ( const QwordsHash_t & h ) override
	{
		if ( m_fIDF<0.0f )
			m_fIDF = h.at ( m_tQword.m_sWord ).m_fIDF;
	}
};

static inline bool IsHitLess ( const ExtHit_t & a, const ExtHit_t & b )	// src/searchnode.cpp:2611-2615
{
	return ( a.m_uHitpos<b.m_uHitpos ) || ( a.m_uHitpos==b.m_uHitpos && a.m_uQuerypos<=b.m_uQuerypos );
}

static void SortHitsReverse ( std::vector<ExtHit_t> & d, size_t iFrom )
{
	// CmpAndHitReverse_fn, src/searchnode.cpp:2618-2624; keys are unique so any sort gives the same order
	std::sort ( d.begin()+iFrom, d.end(), [] ( const ExtHit_t & a, const ExtHit_t & b )
	{
		return ( a.m_uHitpos<b.m_uHitpos ) || ( a.m_uHitpos==b.m_uHitpos && a.m_uQuerypos>b.m_uQuerypos );
	});
}

struct TwoferNode_c : Node_c
{
	std::unique_ptr<Node_c> m_pLeft, m_pRight;
	ExtDoc_t m_tL { INVALID_ROWID, 0, 0.0f }, m_tR { INVALID_ROWID, 0, 0.0f };
	bool m_bHasL = false, m_bHasR = false;		// a pending (unconsumed) doc on that side
	bool m_bEofL = false, m_bEofR = false;
	bool m_bCurL = false, m_bCurR = false;		// which children sit on the doc just returned
	WORD m_uNodePosL = 0, m_uNodePosR = 0;
	std::vector<ExtHit_t> m_dTmpL, m_dTmpR;

	bool PullL()	{ if ( m_bHasL ) return true; if ( m_bEofL ) return false; m_bHasL = m_pLeft->Next ( m_tL ); m_bEofL = !m_bHasL; return m_bHasL; }
	bool PullR()	{ if ( m_bHasR ) return true; if ( m_bEofR ) return false; m_bHasR = m_pRight->Next ( m_tR ); m_bEofR = !m_bHasR; return m_bHasR; }

	int GetQwords ( QwordsHash_t & h ) override
	{
		int iMax1 = m_pLeft->GetQwords ( h );
		int iMax2 = m_pRight->GetQwords ( h );
		return std::max ( iMax1, iMax2 );
	}
	void SetQwordsIDF ( const QwordsHash_t & h ) override	{ m_pLeft->SetQwordsIDF ( h ); m_pRight->SetQwordsIDF ( h ); }

	/// merge by (hitpos, qpos<=), left wins ties: ExtAnd_c::CollectHits :2627-2705, ExtOr_c::CollectHits :3516-3545
	void MergeChildHits ( std::vector<ExtHit_t> & dHits, bool bLeft, bool bRight, bool bSetNodePos )
	{
		m_dTmpL.clear(); m_dTmpR.clear();
		if ( bLeft ) m_pLeft->CollectHits ( m_dTmpL );
		if ( bRight ) m_pRight->CollectHits ( m_dTmpR );
		size_t iFrom = dHits.size();
		size_t l = 0, r = 0;
		while ( l<m_dTmpL.size() || r<m_dTmpR.size() )
		{
			bool bTakeL = r>=m_dTmpR.size() || ( l<m_dTmpL.size() && IsHitLess ( m_dTmpL[l], m_dTmpR[r] ) );
			ExtHit_t t = bTakeL ? m_dTmpL[l++] : m_dTmpR[r++];
			if ( bSetNodePos )
			{
				WORD uPos = bTakeL ? m_uNodePosL : m_uNodePosR;
				if ( uPos )
					t.m_uNodepos = uPos;
			}
			dHits.push_back ( t );
		}
		if ( m_bQPosReverse )
			SortHitsReverse ( dHits, iFrom );
	}
};

/// ExtAnd_c, src/searchnode.cpp:2570-2608
struct AndNode_c : TwoferNode_c
{
	bool Next ( ExtDoc_t & tDoc ) override
	{
		while ( true )
		{
			if ( !m_bHasL )
			{
				if ( m_bHasR ) m_pLeft->HintRowID ( m_tR.m_tRowID );	// WarmupDocs(L,R,left), :127-142
				if ( !PullL() ) return false;
			}
			if ( !m_bHasR )
			{
				m_pRight->HintRowID ( m_tL.m_tRowID );
				if ( !PullR() ) return false;
			}
			if ( m_tL.m_tRowID==m_tR.m_tRowID )
			{
				tDoc.m_tRowID = m_tL.m_tRowID;
				tDoc.m_uDocFields = m_tL.m_uDocFields | m_tR.m_uDocFields;
				tDoc.m_fTFIDF = m_tL.m_fTFIDF + m_tR.m_fTFIDF;
				m_bHasL = m_bHasR = false;
				return true;
			}
			if ( m_tL.m_tRowID<m_tR.m_tRowID ) m_bHasL = false; else m_bHasR = false;
		}
	}
	void HintRowID ( RowID_t t ) override	{ m_pLeft->HintRowID ( t ); m_pRight->HintRowID ( t ); }
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override	{ MergeChildHits ( dHits, true, true, true ); }
};

/// ExtOr_c, src/searchnode.cpp:3465-3551
struct OrNode_c : TwoferNode_c
{
	bool Next ( ExtDoc_t & tDoc ) override
	{
		bool bL = PullL(), bR = PullR();
		if ( !bL && !bR )
			return false;
		if ( bL && bR && m_tL.m_tRowID==m_tR.m_tRowID )
		{
			tDoc = m_tL;
			tDoc.m_uDocFields = m_tL.m_uDocFields | m_tR.m_uDocFields;
			tDoc.m_fTFIDF = m_tL.m_fTFIDF + m_tR.m_fTFIDF;
			m_bHasL = m_bHasR = false;
			m_bCurL = m_bCurR = true;
		} else if ( bL && ( !bR || m_tL.m_tRowID<m_tR.m_tRowID ) )
		{
			tDoc = m_tL;
			m_bHasL = false;
			m_bCurL = true; m_bCurR = false;
		} else
		{
			tDoc = m_tR;
			m_bHasR = false;
			m_bCurL = false; m_bCurR = true;
		}
		return true;
	}
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override	{ MergeChildHits ( dHits, m_bCurL, m_bCurR, false ); }
};

/// ExtMaybe_c, src/searchnode.cpp:3565-3604: left docs; right only adds weight and hits
struct MaybeNode_c : OrNode_c
{
	bool Next ( ExtDoc_t & tDoc ) override
	{
		if ( !PullL() )
			return false;
		while ( PullR() && m_tR.m_tRowID<m_tL.m_tRowID )
			m_bHasR = false;
		tDoc = m_tL;
		m_bCurL = true; m_bCurR = false;
		if ( m_bHasR && m_tR.m_tRowID==m_tL.m_tRowID )
		{
			tDoc.m_uDocFields = m_tL.m_uDocFields | m_tR.m_uDocFields;
			tDoc.m_fTFIDF = m_tL.m_fTFIDF + m_tR.m_fTFIDF;
			m_bHasR = false;
			m_bCurR = true;
		}
		m_bHasL = false;
		return true;
	}
};

/// ExtAndNot_c, src/searchnode.cpp:3618-3711: the right side never yields hits
struct AndNotNode_c : TwoferNode_c
{
	bool Next ( ExtDoc_t & tDoc ) override
	{
		while ( true )
		{
			if ( !PullL() )
				return false;
			while ( PullR() && m_tR.m_tRowID<m_tL.m_tRowID )
				m_bHasR = false;
			m_bHasL = false;
			if ( m_bHasR && m_tR.m_tRowID==m_tL.m_tRowID )
			{
				m_bHasR = false;
				continue;
			}
			tDoc = m_tL;
			return true;
		}
	}
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override	{ m_pLeft->CollectHits ( dHits ); }
};

/// ExtMultiAnd_T, src/searchnode.cpp:2716-3260
struct MultiAndNode_c : Node_c
{
	struct NodeInfo_t
	{
		Qword_t m_tQword;
		DWORD m_uQueriedFields;
		int m_iAtomPos;
		WORD m_uNodepos;
		bool m_bNotWeighted;
		float m_fIDF;
		RowID_t m_tRowID;
		int m_iWordIdx;
		bool FitsFields() const	{ return ( m_tQword.m_uFields & m_uQueriedFields )!=0; }
	};
	std::vector<NodeInfo_t> m_dNodes;
	bool m_bFirst = true;
	bool m_bUseBM25 = true;
	int m_iNodesSet = 0;
	std::vector<uint64_t> m_dCurHitlistPos;
	RowID_t m_tCurRowID = INVALID_ROWID;

	void Finalize()
	{
		RefSort ( m_dNodes, [] ( const NodeInfo_t & a, const NodeInfo_t & b ) { return a.m_tQword.OrderDocs()<b.m_tQword.OrderDocs(); } );	// :2791
		m_dCurHitlistPos.resize ( m_dNodes.size() );
	}
	RowID_t Advance ( int iNode )	// :2845-2855
	{
		NodeInfo_t & t = m_dNodes[iNode];
		do { t.m_tQword.ReadNext(); t.m_tRowID = t.m_tQword.m_tRowID; } while ( t.m_tRowID!=INVALID_ROWID && !t.FitsFields() );
		return t.m_tRowID;
	}
	RowID_t Advance ( int iNode, RowID_t tRowID )	// :2859-2870
	{
		NodeInfo_t & t = m_dNodes[iNode];
		if ( tRowID==t.m_tRowID )
			return tRowID;
		t.m_tRowID = t.m_tQword.AdvanceTo ( tRowID );
		while ( t.m_tRowID!=INVALID_ROWID && !t.FitsFields() )
		{
			t.m_tQword.ReadNext();
			t.m_tRowID = t.m_tQword.m_tRowID;
		}
		return t.m_tRowID;
	}
	bool AdvanceQwords()	// :2874-2899
	{
		RowID_t tMax = m_dNodes[0].m_tRowID;
		for ( int i=1; i<(int)m_dNodes.size(); i++ )
		{
			NodeInfo_t & tCur = m_dNodes[i];
			if ( tCur.m_tRowID==tMax )
				continue;
			Advance ( i, tMax );
			if ( tCur.m_tRowID==INVALID_ROWID )
				return false;
			else if ( tCur.m_tRowID>tMax )
			{
				if ( Advance ( 0, tCur.m_tRowID )==INVALID_ROWID )
					return false;
				tMax = m_dNodes[0].m_tRowID;
				i = 0;
			}
		}
		return true;
	}
	bool Next ( ExtDoc_t & tDoc ) override
	{
		if ( m_bFirst )
		{
			if ( m_iNodesSet!=(int)m_dNodes.size() || !m_dNodes[0].m_tQword.m_iDocs )
				return false;
			for ( auto & n : m_dNodes )
				n.m_tRowID = INVALID_ROWID;
			Advance(0);
			m_bFirst = false;
		} else if ( m_dNodes[0].m_tRowID!=INVALID_ROWID )
			Advance(0);
		if ( m_dNodes[0].m_tRowID==INVALID_ROWID )
			return false;
		if ( !AdvanceQwords() )
		{
			m_dNodes[0].m_tRowID = INVALID_ROWID;
			return false;
		}
		tDoc.m_tRowID = m_dNodes[0].m_tRowID;
		DWORD uMask = 0;
		float fTFIDF = 0.0f;
		for ( size_t i=0; i<m_dNodes.size(); ++i )
		{
			const NodeInfo_t & n = m_dNodes[i];
			uMask |= n.m_tQword.m_uFields & n.m_uQueriedFields;		// :2810-2817
			if ( m_bUseBM25 )
				fTFIDF += float(n.m_tQword.m_uMatchHits) / float(n.m_tQword.m_uMatchHits+SPH_BM25_K1) * n.m_fIDF;	// :2821-2832
			m_dCurHitlistPos[i] = n.m_tQword.m_iHitlistPos;
		}
		tDoc.m_uDocFields = uMask;
		tDoc.m_fTFIDF = fTFIDF;
		m_tCurRowID = tDoc.m_tRowID;
		return true;
	}
	void HintRowID ( RowID_t tRowID ) override	// :3318-3331
	{
		if ( !m_dNodes[0].m_tQword.m_iDocs )
			return;
		if ( m_bFirst || ( m_dNodes[0].m_tRowID!=INVALID_ROWID && tRowID>m_dNodes[0].m_tRowID ) )
		{
			if ( m_bFirst && m_iNodesSet!=(int)m_dNodes.size() )
				return;
			// NB: the reference's Advance(0,rowid) positions node 0 ON a doc; our Next() always steps node 0 first,
			// so only use the skiplist part of the hint here (it never lands on a doc)
			m_dNodes[0].m_tQword.HintRowID ( tRowID );
		}
	}
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override	// MergeHits2/3/N :3098-3181 -- all orders equal (hitpos, qpos)
	{
		size_t iFrom = dHits.size();
		struct Stream_t { Hitpos_t m_uHit; int m_iNode; };
		std::vector<Stream_t> dStreams;
		for ( size_t i=0; i<m_dNodes.size(); ++i )
		{
			m_dNodes[i].m_tQword.SeekHitlist ( m_dCurHitlistPos[i] );
			dStreams.push_back ( { m_dNodes[i].m_tQword.GetNextHit(), (int)i } );
		}
		while ( true )
		{
			int iBest = -1;
			for ( size_t s=0; s<dStreams.size(); ++s )
			{
				if ( dStreams[s].m_uHit==EMPTY_HIT )
					continue;
				if ( iBest<0 ) { iBest = (int)s; continue; }
				const NodeInfo_t & a = m_dNodes[dStreams[s].m_iNode];
				const NodeInfo_t & b = m_dNodes[dStreams[iBest].m_iNode];
				// strictly-better test keeps the earlier stream on full ties, like IsHitLess(L,R) with <=
				if ( dStreams[s].m_uHit<dStreams[iBest].m_uHit || ( dStreams[s].m_uHit==dStreams[iBest].m_uHit && a.m_iAtomPos<b.m_iAtomPos ) )
					iBest = (int)s;
			}
			if ( iBest<0 )
				break;
			NodeInfo_t & n = m_dNodes[dStreams[iBest].m_iNode];
			Hitpos_t uHit = dStreams[iBest].m_uHit;
			int iField = HitField ( uHit );
			if ( iField<32 && ( n.m_uQueriedFields & ( 1u<<iField ) ) )
				dHits.push_back ( { m_tCurRowID, uHit, (WORD)n.m_iAtomPos, n.m_uNodepos, 1, 1, 1, 0 } );
			dStreams[iBest].m_uHit = n.m_tQword.GetNextHit();
		}
		if ( m_bQPosReverse )
			SortHitsReverse ( dHits, iFrom );
	}
	int GetQwords ( QwordsHash_t & h ) override
	{
		int iMax = -1;
		for ( auto & n : m_dNodes )
			iMax = std::max ( iMax, RegisterQword ( n.m_tQword, n.m_bNotWeighted, n.m_fIDF, h, n.m_iWordIdx ) );
		return iMax;
	}
	void SetQwordsIDF ( const QwordsHash_t & h ) override
	{
		for ( auto & n : m_dNodes )
			if ( n.m_fIDF<0.0f )
				n.m_fIDF = h.at ( n.m_tQword.m_sWord ).m_fIDF;
	}
};

/// FSMphrase_c, src/searchnode.cpp:3884-3953
struct FSMphrase_c
{
	struct State_t { int m_iTagQword; DWORD m_uExpHitposWithField; };
	std::vector<int> m_dAtomPos;
	std::vector<int> m_dQposDelta;
	std::vector<State_t> m_dStates;

	void Init ( const std::vector<int> & dAtomPos, int )
	{
		m_dAtomPos = dAtomPos;
		m_dQposDelta.assign ( m_dAtomPos.back()-m_dAtomPos[0]+1, -INT_MAX );
		for ( size_t i=1; i<m_dAtomPos.size(); ++i )
			m_dQposDelta [ m_dAtomPos[i-1]-m_dAtomPos[0] ] = m_dAtomPos[i]-m_dAtomPos[i-1];
	}
	void ResetFSM()	{ m_dStates.clear(); }
	bool HitFSM ( const ExtHit_t * pHit, std::vector<ExtHit_t> & dHits )
	{
		DWORD uHitposWithField = HitPosWithField ( pHit->m_uHitpos );
		if ( pHit->m_uQuerypos==m_dAtomPos[0] )
			m_dStates.push_back ( { 0, uHitposWithField + (DWORD)m_dQposDelta[0] } );
		for ( int i=(int)m_dStates.size()-1; i>=0; i-- )
		{
			if ( m_dStates[i].m_uExpHitposWithField<uHitposWithField )
			{
				m_dStates[i] = m_dStates.back();	// RemoveFast
				m_dStates.pop_back();
				continue;
			}
			if ( m_dStates[i].m_uExpHitposWithField==uHitposWithField && m_dAtomPos [ m_dStates[i].m_iTagQword+1 ]==pHit->m_uQuerypos )
			{
				m_dStates[i].m_iTagQword++;
				m_dStates[i].m_uExpHitposWithField = uHitposWithField + (DWORD)m_dQposDelta [ pHit->m_uQuerypos-m_dAtomPos[0] ];
			}
			if ( m_dStates[i].m_iTagQword==(int)m_dAtomPos.size()-1 )
			{
				DWORD uSpanlen = (DWORD)( m_dAtomPos.back()-m_dAtomPos[0] );
				ExtHit_t t;
				t.m_tRowID = pHit->m_tRowID;
				t.m_uHitpos = uHitposWithField-uSpanlen;
				t.m_uQuerypos = (WORD)m_dAtomPos[0];
				t.m_uNodepos = 0;
				t.m_uMatchlen = t.m_uSpanlen = (WORD)( uSpanlen+1 );
				t.m_uWeight = (DWORD)m_dAtomPos.size();
				t.m_uQposMask = 0;
				dHits.push_back ( t );
				ResetFSM();
				return true;
			}
		}
		return false;
	}
};

/// FSMproximity_c, src/searchnode.cpp:3958-4075
struct FSMproximity_c
{
	int m_iMaxDistance = 0;
	DWORD m_uWordsExpected = 0, m_uMinQpos = 0, m_uQLen = 0, m_uExpPos = 0, m_uWords = 0;
	int m_iMinQindex = -1;
	std::vector<DWORD> m_dProx;
	std::vector<int> m_dDeltas;

	void Init ( const std::vector<int> & dAtomPos, int iOpArg )
	{
		m_iMaxDistance = iOpArg;
		m_uWordsExpected = (DWORD)dAtomPos.size();
		m_uMinQpos = (DWORD)dAtomPos[0];
		m_uQLen = (DWORD)( dAtomPos.back()-dAtomPos[0] );
		m_dProx.resize ( m_uQLen+1 );
		m_dDeltas.resize ( m_uQLen+1 );
	}
	void ResetFSM()
	{
		m_uExpPos = 0; m_uWords = 0; m_iMinQindex = -1;
		std::fill ( m_dProx.begin(), m_dProx.end(), UINT_MAX );
	}
	bool HitFSM ( const ExtHit_t * pHit, std::vector<ExtHit_t> & dHits )
	{
		int iQindex = (int)pHit->m_uQuerypos - (int)m_uMinQpos;
		DWORD uHitposWithField = HitPosWithField ( pHit->m_uHitpos );
		if ( m_dProx[iQindex]==UINT_MAX )
			m_uWords++;
		m_dProx[iQindex] = uHitposWithField;
		if ( uHitposWithField>=m_uExpPos || iQindex==m_iMinQindex )
		{
			m_iMinQindex = iQindex;
			int iMinPos = (int)( uHitposWithField-m_uQLen-(DWORD)m_iMaxDistance );
			for ( size_t i=0; i<m_dProx.size(); ++i )
				if ( m_dProx[i]!=UINT_MAX )
				{
					if ( (int)m_dProx[i]<=iMinPos )
					{
						m_dProx[i] = UINT_MAX;
						m_uWords--;
						continue;
					}
					if ( m_dProx[i]<uHitposWithField )
					{
						m_iMinQindex = (int)i;
						uHitposWithField = m_dProx[i];
					}
				}
			m_uExpPos = m_dProx[m_iMinQindex] + m_uQLen + (DWORD)m_iMaxDistance;
		}
		if ( m_uWords!=m_uWordsExpected )
			return false;

		DWORD uMax = 0;
		for ( size_t i=0; i<m_dProx.size(); ++i )
			if ( m_dProx[i]!=UINT_MAX )
			{
				m_dDeltas[i] = (int)( m_dProx[i]-(DWORD)i );
				uMax = std::max ( uMax, m_dProx[i] );
			} else
				m_dDeltas[i] = INT_MAX;
		std::sort ( m_dDeltas.begin(), m_dDeltas.end() );

		DWORD uCurWeight = 0, uWeight = 0;
		int iLast = -INT_MAX;
		for ( size_t i=0; i<m_dDeltas.size() && m_dDeltas[i]!=INT_MAX; ++i )
		{
			if ( m_dDeltas[i]==iLast )
				uCurWeight++;
			else
			{
				uWeight += uCurWeight ? ( 1+uCurWeight ) : 0;
				uCurWeight = 0;
			}
			iLast = m_dDeltas[i];
		}
		uWeight += uCurWeight ? ( 1+uCurWeight ) : 0;
		if ( !uWeight )
			uWeight = 1;

		ExtHit_t t;
		t.m_tRowID = pHit->m_tRowID;
		t.m_uHitpos = m_dProx[m_iMinQindex];
		t.m_uQuerypos = (WORD)m_uMinQpos;
		t.m_uNodepos = 0;
		t.m_uSpanlen = t.m_uMatchlen = (WORD)( uMax-m_dProx[m_iMinQindex]+1 );
		t.m_uWeight = uWeight;
		t.m_uQposMask = 0;
		dHits.push_back ( t );

		m_dProx[m_iMinQindex] = UINT_MAX;
		m_iMinQindex = -1;
		m_uWords--;
		m_uExpPos = 0;
		return true;
	}
};

/// FSMmultinear_c, src/searchnode.cpp:680-720, 4080-4316: `a NEAR/N b NEAR/N c` (children are nodes: keywords, phrases, groups, other
/// NEAR nodes). Hits arrive ordered by position; m_uNodepos tells which child a hit belongs to. A chain grows while every next hit
/// starts within N positions after the previous one ended and belongs to a child not seen in the chain yet; two children: every
/// adjacent pair is emitted (chains may overlap), more children: the chain is emitted when all of them are in and then starts over.
/// Quirks kept: ResetFSM() leaves the smallest query position seen so far (n-way chains) and the node-position list alone.
struct FSMmultinear_c
{
	int m_iNear = 1;
	DWORD m_uPrelastP = 0, m_uPrelastML = 0, m_uPrelastSL = 0, m_uPrelastW = 0;
	DWORD m_uLastP = 0, m_uLastML = 0, m_uLastSL = 0, m_uLastW = 0;
	DWORD m_uWordsExpected = 0, m_uWeight = 0, m_uFirstHit = 0;
	WORD m_uFirstNpos = 0, m_uFirstQpos = 65535;
	std::vector<WORD> m_dNpos;
	std::vector<ExtHit_t> m_dRing;
	int m_iRing = 0;
	bool m_bTwofer = false, m_bQposMask = false;

	void InitNear ( int nNodes, int iNear, bool bQposMask )
	{
		m_iNear = iNear;
		m_uWordsExpected = (DWORD)nNodes;
		m_bQposMask = bQposMask;
		m_bTwofer = ( nNodes==2 );
		if ( !m_bTwofer )
			m_dRing.resize ( nNodes );
	}
	void ResetFSM()				{ m_iRing = 0; m_uLastP = 0; m_uPrelastP = 0; }
	int RingTail() const		{ return ( m_iRing + (int)m_dNpos.size() - 1 ) % (int)m_uWordsExpected; }
	void Add2Ring ( const ExtHit_t * pHit )	{ if ( !m_bTwofer ) m_dRing[RingTail()] = *pHit; }
	void ShiftRing()			{ if ( ++m_iRing==(int)m_uWordsExpected ) m_iRing = 0; }
	void StartChain ( const ExtHit_t * pHit, DWORD uPos )
	{
		m_uFirstHit = m_uLastP = uPos;
		m_uLastML = pHit->m_uMatchlen;
		m_uLastSL = pHit->m_uSpanlen;
		m_uWeight = m_uLastW = pHit->m_uWeight;
	}

	bool HitFSM ( const ExtHit_t * pHit, std::vector<ExtHit_t> & dHits )
	{
		const DWORD uPos = HitPosWithField ( pHit->m_uHitpos );
		const WORD uNpos = pHit->m_uNodepos, uQpos = pHit->m_uQuerypos;

		// a second hit at the position of the last one (an OR child, `a NEAR/2 a`...), :4103-4134
		if ( m_uLastP==uPos )
		{
			if ( m_bTwofer && uNpos<m_uFirstNpos )
			{
				m_uFirstQpos = uQpos;	// keep the leftmost child of the query
				m_uFirstNpos = uNpos;
				return false;
			}
			if ( !m_bTwofer && uNpos<m_dRing[RingTail()].m_uNodepos )
			{
				if ( !std::binary_search ( m_dNpos.begin(), m_dNpos.end(), uNpos ) )
				{
					auto it = std::lower_bound ( m_dNpos.begin(), m_dNpos.end(), m_dRing[RingTail()].m_uNodepos );
					*it = uNpos;
					std::sort ( m_dNpos.begin(), m_dNpos.end() );
					m_dRing[RingTail()].m_uNodepos = uNpos;
					m_dRing[RingTail()].m_uQuerypos = uQpos;
				}
				return false;
			}
			if ( m_uPrelastP && m_uLastML<pHit->m_uMatchlen )
			{
				// the last hit was a part of this longer one: step back to the one before it
				m_uLastML = m_uPrelastML;
				m_uLastSL = m_uPrelastSL;
				m_uFirstHit = m_uLastP = m_uPrelastP;
				m_uWeight = m_uWeight - m_uLastW + m_uPrelastW;
			} else
				return false;
		}

		// too far from the previous hit (or no previous hit): a new chain starts here, :4137-4154
		if ( m_uLastP==0 || ( m_uLastP + m_uLastML + m_iNear )<=uPos )
		{
			StartChain ( pHit, uPos );
			if ( m_bTwofer )
			{
				m_uFirstQpos = uQpos;
				m_uFirstNpos = uNpos;
			} else
			{
				m_dNpos.assign ( 1, uNpos );
				Add2Ring ( pHit );
			}
			return false;
		}

		if ( m_bTwofer )
		{
			// overlapping hits of different length: restart from the new one, :4160-4172
			if ( ( m_uFirstHit + m_uLastML )>uPos && ( m_uFirstHit + m_uLastML )<( uPos + pHit->m_uMatchlen ) && m_uLastML!=pHit->m_uMatchlen )
			{
				StartChain ( pHit, uPos );
				m_uFirstQpos = uQpos;
				m_uFirstNpos = uNpos;
				return false;
			}
			// the same child again: it becomes the head of the chain, :4173-4190
			if ( uNpos==m_uFirstNpos )
			{
				if ( m_uLastP<uPos )
				{
					m_uPrelastML = m_uLastML;
					m_uPrelastSL = m_uLastSL;
					m_uPrelastP = m_uLastP;
					m_uPrelastW = pHit->m_uWeight;
					m_uFirstHit = m_uLastP = uPos;
					m_uLastML = pHit->m_uMatchlen;
					m_uLastSL = pHit->m_uSpanlen;
					m_uWeight = m_uLastW = m_uPrelastW;
					m_uFirstQpos = uQpos;
					m_uFirstNpos = uNpos;
				}
				return false;
			}
		} else
		{
			// n-way: the chain keeps a sorted list of the children it holds and a ring of their hits, :4193-4247
			if ( uNpos<m_dNpos.front() )
			{
				m_uFirstQpos = std::min ( m_uFirstQpos, uQpos );
				m_dNpos.insert ( m_dNpos.begin(), uNpos );
			} else if ( uNpos>m_dNpos.back() )
			{
				m_uFirstQpos = std::min ( m_uFirstQpos, uQpos );
				m_dNpos.push_back ( uNpos );
			} else if ( uNpos!=m_dNpos.front() && uNpos!=m_dNpos.back() )
			{
				int iEnd = (int)m_dNpos.size(), iStart = 0;
				while ( iEnd-iStart>1 )
				{
					const int iMid = ( iStart+iEnd )/2;
					if ( uNpos==m_dNpos[iMid] )
					{
						const ExtHit_t & tHead = m_dRing[m_iRing];
						if ( uNpos==tHead.m_uNodepos )
						{
							// the child at the head of the chain again: drop the head
							m_uWeight -= tHead.m_uWeight;
							m_uFirstHit = HitPosWithField ( tHead.m_uHitpos );
							ShiftRing();
						} else if ( uNpos==m_dRing[RingTail()].m_uNodepos )
							m_uWeight -= m_dRing[RingTail()].m_uWeight;	// the child at the tail again: the new hit replaces it
						else
							return false;
					}
					if ( uNpos<m_dNpos[iMid] )
						iEnd = iMid;
					else
						iStart = iMid;
				}
				m_dNpos.insert ( m_dNpos.begin()+iEnd, uNpos );
				m_uFirstQpos = std::min ( m_uFirstQpos, uQpos );
			} else if ( uNpos==m_dRing[m_iRing].m_uNodepos )
			{
				m_uWeight -= m_dRing[m_iRing].m_uWeight;
				m_uFirstHit = HitPosWithField ( m_dRing[m_iRing].m_uHitpos );
				ShiftRing();
			} else if ( uNpos==m_dRing[RingTail()].m_uNodepos )
				m_uWeight -= m_dRing[RingTail()].m_uWeight;
			else
				return false;
		}

		m_uWeight += pHit->m_uWeight;
		m_uLastML = pHit->m_uMatchlen;
		m_uLastSL = pHit->m_uSpanlen;
		Add2Ring ( pHit );

		// the whole chain is there: emit it, :4254-4286
		if ( m_bTwofer || m_uWordsExpected==(DWORD)m_dNpos.size() )
		{
			ExtHit_t t;
			t.m_tRowID = pHit->m_tRowID;
			t.m_uHitpos = m_uFirstHit;
			t.m_uNodepos = 0;
			t.m_uMatchlen = (WORD)( uPos - m_uFirstHit + m_uLastML );
			t.m_uWeight = m_uWeight;
			m_uPrelastP = 0;
			t.m_uQuerypos = std::min ( m_uFirstQpos, pHit->m_uQuerypos );
			if ( m_bTwofer )
			{
				// two children may overlap: the chain shifts instead of starting over
				t.m_uSpanlen = 2;
				t.m_uQposMask = 1u<<( std::max ( m_uFirstQpos, pHit->m_uQuerypos ) - t.m_uQuerypos );
				m_uFirstHit = m_uLastP = uPos;
				m_uWeight = pHit->m_uWeight;
				m_uFirstQpos = pHit->m_uQuerypos;
			} else
			{
				t.m_uSpanlen = (WORD)m_dNpos.size();
				t.m_uQposMask = 0;
				m_uLastP = 0;
				if ( m_bQposMask && t.m_uSpanlen>1 )
					for ( WORD uN : m_dNpos )
						t.m_uQposMask |= 1u<<( uN - t.m_uQuerypos );
			}
			dHits.push_back ( t );
			return true;
		}
		m_uLastP = uPos;
		return false;
	}
};

/// ExtNWay_T<FSM>, src/searchnode.cpp:3767-3848
template<typename FSM>
struct NWayNode_c : Node_c, FSM
{
	std::unique_ptr<Node_c> m_pNode;
	std::vector<ExtHit_t> m_dRaw, m_dMyHits;

	bool Next ( ExtDoc_t & tDoc ) override
	{
		ExtDoc_t tCand;
		while ( m_pNode->Next ( tCand ) )
		{
			m_dRaw.clear();
			m_pNode->CollectHits ( m_dRaw );
			m_dMyHits.clear();
			FSM::ResetFSM();
			bool bEmitted = false;
			for ( const ExtHit_t & tHit : m_dRaw )
				if ( FSM::HitFSM ( &tHit, m_dMyHits ) && !bEmitted )
				{
					bEmitted = true;
					tDoc.m_tRowID = tHit.m_tRowID;
					tDoc.m_uDocFields = 1u<<( HitField ( tHit.m_uHitpos ) & 31 );
					tDoc.m_fTFIDF = tCand.m_fTFIDF;
				}
			if ( bEmitted )
				return true;
		}
		return false;
	}
	void HintRowID ( RowID_t t ) override	{ m_pNode->HintRowID ( t ); }
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override	{ dHits.insert ( dHits.end(), m_dMyHits.begin(), m_dMyHits.end() ); }
	int GetQwords ( QwordsHash_t & h ) override			{ return m_pNode->GetQwords ( h ); }
	void SetQwordsIDF ( const QwordsHash_t & h ) override	{ m_pNode->SetQwordsIDF ( h ); }
};

//////////////////////////////////////////////////////////////////////////
// factory: ExtNode_i::Create, src/searchnode.cpp:1599-1811
//////////////////////////////////////////////////////////////////////////

/// ExtQuorum_c, src/searchnode.cpp:4319-4650: a document matches when at least m_iThresh of the keywords occur in it (a keyword
/// repeated in the query counts as often as it is repeated, but only up to its number of hits in the document).
/// The children vector keeps the reference's dynamic order: keywords sorted by query position, an exhausted keyword is removed
/// with RemoveFast (the last one takes its slot), and TF*IDF is summed in that order.
struct QuorumNode_c : Node_c
{
	struct Child_t
	{
		std::shared_ptr<TermNode_c> m_pTerm;
		int m_iCount = 1;
		ExtDoc_t m_tDoc { INVALID_ROWID, 0, 0.0f };
		bool m_bHas = false;
	};
	std::vector<Child_t> m_dInitial, m_dChildren;
	int m_iThresh = 1;
	bool m_bHasDupes = false;
	bool m_bWarm = false;
	int m_iQuorumLeft = 0;
	std::vector<ExtHit_t> m_dCurHits, m_dTmp;

	/// ctor, :4342-4404: fold repeated keywords into one child with a count (the first occurrence stays), sort back by query position
	void Init ( std::vector<TermNode_c*> & dTerms, int iThresh )
	{
		m_iThresh = std::max ( iThresh, 1 );
		m_iAtomPos = dTerms[0]->m_iAtomPos;
		for ( size_t i=0; i<dTerms.size(); ++i )
		{
			size_t iParent = m_dInitial.size();
			for ( size_t j=0; j<m_dInitial.size(); ++j )
				if ( m_dInitial[j].m_pTerm->m_tQword.m_sWord==dTerms[i]->m_tQword.m_sWord )
					iParent = j;
			if ( iParent<m_dInitial.size() )
			{
				m_dInitial[iParent].m_iCount++;
				m_bHasDupes = true;
				delete dTerms[i];
			} else
			{
				Child_t t;
				t.m_pTerm.reset ( dTerms[i] );
				m_dInitial.push_back ( t );
			}
		}
		std::sort ( m_dInitial.begin(), m_dInitial.end(), [] ( const Child_t & a, const Child_t & b ) { return a.m_pTerm->m_iAtomPos<b.m_pTerm->m_iAtomPos; } );
		m_dChildren = m_dInitial;
	}

	void RemoveFast ( size_t i )
	{
		m_dChildren[i] = m_dChildren.back();
		m_dChildren.pop_back();
	}
	bool Pull ( Child_t & t )
	{
		t.m_bHas = t.m_pTerm->Next ( t.m_tDoc );
		return t.m_bHas;
	}
	/// CountQuorum, :4574-4595
	int CountQuorum ( bool bFixDupes )
	{
		if ( !m_bHasDupes )
			return (int)m_dChildren.size();
		int iSum = 0;
		bool bHasDupes = false;
		for ( auto & t : m_dChildren )
		{
			iSum += t.m_iCount;
			bHasDupes |= ( t.m_iCount>1 );
		}
		m_bHasDupes = bFixDupes ? bHasDupes : m_bHasDupes;
		return iSum;
	}
	/// CollectMatchingHits, :4602-4650; the hits of every keyword on the row end up in m_dCurHits
	bool CollectMatchingHits ( RowID_t tRowID )
	{
		m_dCurHits.clear();
		int iQuorum = 0;
		bool bCounting = m_bHasDupes;
		for ( auto & t : m_dChildren )
		{
			if ( !t.m_bHas || t.m_tDoc.m_tRowID!=tRowID )
				continue;
			m_dTmp.clear();
			t.m_pTerm->CollectHits ( m_dTmp );
			if ( bCounting )
			{
				// matched hits count only up to the keyword's repeat count
				iQuorum += std::min<int> ( t.m_iCount, (int)m_dTmp.size() );
				if ( iQuorum>=m_iThresh )
					bCounting = false;
			}
			m_dCurHits.insert ( m_dCurHits.end(), m_dTmp.begin(), m_dTmp.end() );
		}
		if ( m_bHasDupes && iQuorum<m_iThresh )
		{
			m_dCurHits.clear();
			return false;
		}
		// CollectHits: QuorumCmpHitPos_fn, :4543-4565
		std::stable_sort ( m_dCurHits.begin(), m_dCurHits.end(), [] ( const ExtHit_t & a, const ExtHit_t & b )
		{
			DWORD uA = HitPosWithField ( a.m_uHitpos ), uB = HitPosWithField ( b.m_uHitpos );
			return uA<uB || ( uA==uB && a.m_uQuerypos<b.m_uQuerypos );
		});
		return true;
	}

	/// GetDocsChunk, :4466-4541, one document per call
	bool Next ( ExtDoc_t & tDoc ) override
	{
		if ( !m_bWarm )
		{
			m_bWarm = true;
			for ( size_t i=0; i<m_dChildren.size(); ++i )
				if ( !Pull ( m_dChildren[i] ) )
				{
					RemoveFast ( i );
					--i;
				}
			m_iQuorumLeft = CountQuorum ( true );
		}
		while ( m_iQuorumLeft>=m_iThresh )
		{
			ExtDoc_t tCand { INVALID_ROWID, 0, 0.0f };
			int iQuorum = 0;
			for ( auto & t : m_dChildren )
			{
				if ( t.m_tDoc.m_tRowID<tCand.m_tRowID )
				{
					tCand = t.m_tDoc;
					iQuorum = t.m_iCount;
				} else if ( t.m_tDoc.m_tRowID==tCand.m_tRowID )
				{
					tCand.m_uDocFields |= t.m_tDoc.m_uDocFields;
					tCand.m_fTFIDF += t.m_tDoc.m_fTFIDF;
					iQuorum += t.m_iCount;
				}
			}
			const bool bMatch = iQuorum>=m_iThresh && CollectMatchingHits ( tCand.m_tRowID );

			// advance the children that sat on the candidate
			const size_t nBefore = m_dChildren.size();
			for ( size_t i=0; i<m_dChildren.size(); ++i )
			{
				if ( m_dChildren[i].m_tDoc.m_tRowID!=tCand.m_tRowID )
					continue;
				if ( !Pull ( m_dChildren[i] ) )
				{
					RemoveFast ( i );
					--i;
				}
			}
			if ( nBefore!=m_dChildren.size() )
				m_iQuorumLeft = CountQuorum ( false );
			if ( bMatch )
			{
				tDoc = tCand;
				return true;
			}
		}
		return false;
	}
	void HintRowID ( RowID_t tRowID ) override
	{
		for ( auto & t : m_dChildren )
			t.m_pTerm->HintRowID ( tRowID );
	}
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override
	{
		dHits.insert ( dHits.end(), m_dCurHits.begin(), m_dCurHits.end() );
	}
	int GetQwords ( QwordsHash_t & h ) override
	{
		int iMax = -1;
		for ( auto & t : m_dChildren )
			iMax = std::max ( iMax, t.m_pTerm->GetQwords ( h ) );
		return iMax;
	}
	void SetQwordsIDF ( const QwordsHash_t & h ) override
	{
		for ( auto & t : m_dChildren )
			t.m_pTerm->SetQwordsIDF ( h );
	}
};

/// ExtOrder_c, src/searchnode.cpp:4657-4935: `a << b << c`. A document has to hold every child; its hits are walked in position order
/// (the lowest child wins a tie) and two trackers, the longest prefix of the child sequence found so far and the most recent attempt,
/// look for the children in query order inside ONE field; every complete sequence is emitted. The document carries child 0's
/// TF*IDF and field mask only (:4914).
struct OrderNode_c : Node_c
{
	std::vector<std::unique_ptr<Node_c>> m_dChildren;
	std::vector<ExtDoc_t> m_dDoc;
	std::vector<char> m_dHas;
	bool m_bDone = false;
	std::vector<std::vector<ExtHit_t>> m_dChildHits;
	std::vector<ExtHit_t> m_dCurHits;

	bool Pull ( size_t i )
	{
		m_dHas[i] = m_dChildren[i]->Next ( m_dDoc[i] ) ? 1 : 0;
		return m_dHas[i]!=0;
	}
	/// GetMatchingHits, :4734-4829
	bool GetMatchingHits()
	{
		const size_t n = m_dChildren.size();
		std::vector<size_t> dCur ( n, 0 );
		std::vector<ExtHit_t> dLongest, dRecent;
		int iPosLongest = 0, iPosRecent = 0, iField = -1;
		m_dCurHits.clear();
		while ( true )
		{
			// next hit in position order; the lowest child wins ties (:4706-4731)
			int iChild = -1;
			DWORD uMin = UINT_MAX;
			for ( size_t i=0; i<n; ++i )
				if ( dCur[i]<m_dChildHits[i].size() && HitPosWithField ( m_dChildHits[i][dCur[i]].m_uHitpos )<uMin )
				{
					uMin = HitPosWithField ( m_dChildHits[i][dCur[i]].m_uHitpos );
					iChild = (int)i;
				}
			if ( iChild<0 )
				break;
			const ExtHit_t & tHit = m_dChildHits[iChild][dCur[iChild]++];
			const int iHitField = HitField ( tHit.m_uHitpos ), iHitPos = (int)( tHit.m_uHitpos & 0x7FFFFFu );
			if ( iHitField!=iField )
			{
				// another field: both trackers start over; only child 0 can seed (and only then the field is remembered)
				dLongest.clear();
				dRecent.clear();
				if ( iChild==0 )
				{
					dLongest.push_back ( tHit );
					iPosLongest = iHitPos + tHit.m_uSpanlen;
					iField = iHitField;
				}
			} else if ( iChild==(int)dLongest.size() && iHitPos>=iPosLongest )
			{
				dLongest.push_back ( tHit );
				iPosLongest = iHitPos + tHit.m_uSpanlen;
				if ( dLongest.size()==n )
				{
					m_dCurHits.insert ( m_dCurHits.end(), dLongest.begin(), dLongest.end() );
					dLongest.clear();
					dRecent.clear();
					iPosRecent = iPosLongest;
				}
			} else if ( iChild==0 )
			{
				dRecent.assign ( 1, tHit );
				iPosRecent = iHitPos + tHit.m_uSpanlen;
				if ( dLongest.empty() )
				{
					dLongest.push_back ( tHit );
					iPosLongest = iHitPos + tHit.m_uSpanlen;
				}
			} else if ( iChild==(int)dRecent.size() && iHitPos>=iPosRecent )
			{
				dRecent.push_back ( tHit );
				iPosRecent = iHitPos + tHit.m_uSpanlen;
				if ( dRecent.size()==dLongest.size() )
				{
					dLongest.swap ( dRecent );
					dRecent.clear();
					iPosLongest = iPosRecent;
				}
			}
		}
		return !m_dCurHits.empty();
	}
	/// GetDocsChunk, :4832-4929, one document per call
	bool Next ( ExtDoc_t & tDoc ) override
	{
		const size_t n = m_dChildren.size();
		while ( !m_bDone )
		{
			for ( size_t i=0; i<n; ++i )
				if ( !m_dHas[i] && !Pull ( i ) )
				{
					m_bDone = true;
					return false;
				}
			// the next document that holds every child
			RowID_t tRowID = m_dDoc[0].m_tRowID;
			size_t iChild = 1;
			while ( iChild<n )
			{
				while ( m_dDoc[iChild].m_tRowID<tRowID )
					if ( !Pull ( iChild ) )
					{
						m_bDone = true;
						return false;
					}
				if ( m_dDoc[iChild].m_tRowID>tRowID )
				{
					tRowID = m_dDoc[iChild].m_tRowID;
					iChild = 0;
					continue;
				}
				++iChild;
			}
			for ( size_t i=0; i<n; ++i )
			{
				m_dChildHits[i].clear();
				m_dChildren[i]->CollectHits ( m_dChildHits[i] );
			}
			const bool bMatch = GetMatchingHits();
			tDoc = m_dDoc[0];
			m_dHas[0] = 0;	// advance child 0; the others catch up on the next call
			if ( bMatch )
				return true;
		}
		return false;
	}
	void HintRowID ( RowID_t t ) override
	{
		for ( auto & p : m_dChildren )
			p->HintRowID ( t );
	}
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override	{ dHits.insert ( dHits.end(), m_dCurHits.begin(), m_dCurHits.end() ); }
	int GetQwords ( QwordsHash_t & h ) override
	{
		int iMax = -1;
		for ( auto & p : m_dChildren )
			iMax = std::max ( iMax, p->GetQwords ( h ) );
		return iMax;
	}
	void SetQwordsIDF ( const QwordsHash_t & h ) override
	{
		for ( auto & p : m_dChildren )
			p->SetQwordsIDF ( h );
	}
};

/// ExtNotNear_c, src/searchnode.cpp:5325-5478: `must NOTNEAR/N not`. Every MUST document stays; where the NOT child holds the document
/// too, a MUST hit survives when the first NOT hit at or after it starts more than N positions past its end (NOT hits before it do
/// not count), and the document stays when at least one MUST hit survives. It carries the MUST child's TF*IDF and fields.
struct NotNearNode_c : TwoferNode_c
{
	int m_iDist = 1;
	std::vector<ExtHit_t> m_dCurHits;

	bool Next ( ExtDoc_t & tDoc ) override
	{
		while ( PullL() )
		{
			m_pRight->HintRowID ( m_tL.m_tRowID );
			while ( PullR() && m_tR.m_tRowID<m_tL.m_tRowID )
				m_bHasR = false;
			m_dTmpL.clear();
			m_pLeft->CollectHits ( m_dTmpL );
			m_dCurHits.clear();
			bool bMatched = true;
			if ( m_bHasR && m_tR.m_tRowID==m_tL.m_tRowID )
			{
				// FilterHits, :5352-5380
				m_dTmpR.clear();
				m_pRight->CollectHits ( m_dTmpR );
				size_t iNot = 0;
				for ( size_t iMust=0; iMust<m_dTmpL.size(); ++iMust )
				{
					const DWORD uPosMust = HitPosWithField ( m_dTmpL[iMust].m_uHitpos );
					while ( iNot<m_dTmpR.size() && HitPosWithField ( m_dTmpR[iNot].m_uHitpos )<uPosMust )
						++iNot;
					if ( iNot==m_dTmpR.size() )
					{
						// no NOT hit behind this one: it and the rest of the MUST hits stay
						m_dCurHits.insert ( m_dCurHits.end(), m_dTmpL.begin()+iMust, m_dTmpL.end() );
						break;
					}
					// (the field sits in the top byte, so the distance can be added to the position as it is)
					if ( uPosMust + m_dTmpL[iMust].m_uMatchlen - 1 + m_iDist<HitPosWithField ( m_dTmpR[iNot].m_uHitpos ) )
						m_dCurHits.push_back ( m_dTmpL[iMust] );
				}
				bMatched = !m_dCurHits.empty();
				m_bHasR = false;
			} else
				m_dCurHits = m_dTmpL;
			tDoc = m_tL;
			m_bHasL = false;
			if ( bMatched )
				return true;
		}
		return false;
	}
	void HintRowID ( RowID_t t ) override	{ m_pLeft->HintRowID ( t ); }
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override	{ dHits.insert ( dHits.end(), m_dCurHits.begin(), m_dCurHits.end() ); }
};

/// ExtUnit_c, src/searchnode.cpp:4958-5330: SENTENCE / PARAGRAPH = an AND whose hits must not be separated by a hit of the boundary
/// keyword ("\3sentence" / "\3paragraph", written by the indexing side under index_sp); only the hits of matching units go up
struct UnitNode_c : TwoferNode_c
{
	std::unique_ptr<TermNode_c> m_pDot;
	ExtDoc_t m_tDot { INVALID_ROWID, 0, 0.0f };
	bool m_bHasDot = false, m_bEofDot = false;
	std::vector<ExtHit_t> m_dCurHits, m_dDotHits;

	bool PullDot()	{ if ( m_bHasDot ) return true; if ( m_bEofDot ) return false; m_bHasDot = m_pDot->Next ( m_tDot ); m_bEofDot = !m_bHasDot; return m_bHasDot; }

	/// FilterHits, :5082-5163 (positions compared raw, as there)
	bool FilterHits ()
	{
		const std::vector<ExtHit_t> & h1 = m_dTmpL, & h2 = m_dTmpR, & hd = m_dDotHits;
		size_t i1 = 0, i2 = 0, id = 0;
		DWORD uSentenceEnd = hd.empty() ? UINT_MAX : 0;		// no dots in the document: it degenerates into AND
		bool bRegistered = false;
		while ( true )
		{
			if ( uSentenceEnd )
			{
				const bool bValid1 = i1<h1.size() && h1[i1].m_uHitpos<uSentenceEnd;
				const bool bValid2 = i2<h2.size() && h2[i2].m_uHitpos<uSentenceEnd;
				if ( !bValid1 && !bValid2 )
				{
					uSentenceEnd = 0;
					if ( i1<h1.size() && i2<h2.size() )
						continue;	// perhaps more sentences in this document
					break;
				}
				bRegistered = true;
				if ( bValid1 && ( !bValid2 || IsHitLess ( h1[i1], h2[i2] ) ) )
					m_dCurHits.push_back ( h1[i1++] );
				else
					m_dCurHits.push_back ( h2[i2++] );
			} else
			{
				const DWORD uMin = std::min ( h1[i1].m_uHitpos, h2[i2].m_uHitpos );
				const DWORD uMax = std::max ( h1[i1].m_uHitpos, h2[i2].m_uHitpos );
				while ( id<hd.size() && hd[id].m_uHitpos<=uMin )
					++id;
				if ( id>=hd.size() )
				{
					uSentenceEnd = UINT_MAX;	// no more dots past the pair: everything up to the end of the document matches
					continue;
				}
				if ( hd[id].m_uHitpos<uMax )
				{
					// "A dot B": both sides move past this dot
					const DWORD uDot = hd[id].m_uHitpos;
					while ( i1<h1.size() && h1[i1].m_uHitpos<=uDot ) ++i1;
					if ( i1>=h1.size() ) break;
					while ( i2<h2.size() && h2[i2].m_uHitpos<=uDot ) ++i2;
					if ( i2>=h2.size() ) break;
					continue;
				}
				while ( id<hd.size() && hd[id].m_uHitpos<=uMax )
					++id;
				uSentenceEnd = id>=hd.size() ? UINT_MAX : hd[id].m_uHitpos;
			}
		}
		return bRegistered;
	}

	bool Next ( ExtDoc_t & tDoc ) override
	{
		while ( true )
		{
			if ( !m_bHasL )
			{
				if ( m_bHasR ) m_pLeft->HintRowID ( m_tR.m_tRowID );
				if ( !PullL() ) return false;
			}
			if ( !m_bHasR )
			{
				m_pRight->HintRowID ( m_tL.m_tRowID );
				if ( !PullR() ) return false;
			}
			if ( m_tL.m_tRowID!=m_tR.m_tRowID )
			{
				if ( m_tL.m_tRowID<m_tR.m_tRowID ) m_bHasL = false; else m_bHasR = false;
				continue;
			}
			const RowID_t tRowID = m_tL.m_tRowID;
			m_pDot->HintRowID ( tRowID );
			while ( PullDot() && m_tDot.m_tRowID<tRowID )
				m_bHasDot = false;
			m_dTmpL.clear(); m_dTmpR.clear(); m_dDotHits.clear(); m_dCurHits.clear();
			m_pLeft->CollectHits ( m_dTmpL );
			m_pRight->CollectHits ( m_dTmpR );
			if ( m_bHasDot && m_tDot.m_tRowID==tRowID )
			{
				m_pDot->CollectHits ( m_dDotHits );
				m_bHasDot = false;
			}
			tDoc.m_tRowID = tRowID;
			tDoc.m_uDocFields = m_tL.m_uDocFields | m_tR.m_uDocFields;
			tDoc.m_fTFIDF = m_tL.m_fTFIDF + m_tR.m_fTFIDF;
			m_bHasL = m_bHasR = false;
			if ( !m_dTmpL.empty() && !m_dTmpR.empty() && FilterHits() )
				return true;
		}
	}
	void HintRowID ( RowID_t t ) override	{ m_pLeft->HintRowID ( t ); m_pRight->HintRowID ( t ); }
	void CollectHits ( std::vector<ExtHit_t> & dHits ) override	{ dHits.insert ( dHits.end(), m_dCurHits.begin(), m_dCurHits.end() ); }
};

struct Setup_t
{
	const Index_t * m_pIndex;
	const mgpu_query * m_pQuery;
	bool m_bUseBM25;
	int m_iError = MGPU_OK;
};

static Node_c * CreateNode ( int iNode, Setup_t & tSetup );

static TermNode_c * CreateTerm ( const mgpu_xqnode & tNode, int iWord, Setup_t & tSetup )
{
	const mgpu_xqkeyword & tWord = tSetup.m_pQuery->words[iWord];
	TermNode_c * p = new TermNode_c;
	p->m_iTermPos = ( tWord.field_start && tWord.field_end ) ? 3 : tWord.field_start ? 1 : tWord.field_end ? 2 : 0;
	if ( tNode.field_max_pos )
	{
		p->m_iTermPos = 4;
		p->m_iMaxFieldPos = tNode.field_max_pos;
	}
	p->m_tQword.Setup ( tSetup.m_pIndex, tWord.word );
	if ( tSetup.m_pQuery->shard_of_global && tSetup.m_pQuery->word_docs && tSetup.m_pQuery->word_docs[iWord]>=0 )
		p->m_tQword.m_iOrderDocs = tSetup.m_pQuery->word_docs[iWord];
	p->m_tQword.m_sWord = tWord.word;
	p->m_tQword.m_iAtomPos = tWord.atom_pos;
	p->m_tQword.m_fBoost = tWord.boost;
	p->m_tQword.m_bExcluded = tWord.excluded!=0;
	p->m_tQword.m_bExpanded = tWord.expanded!=0;
	p->m_tQword.m_pIndex = tSetup.m_pIndex;
	p->m_iAtomPos = tWord.atom_pos;
	p->m_uQueriedFields = tNode.field_mask;
	p->m_bNotWeighted = tNode.not_weighted!=0;
	p->m_bUseBM25 = tSetup.m_bUseBM25;
	p->m_iWordIdx = iWord;
	return p;
}

template<typename FSM>
static Node_c * CreateMultiNode ( const mgpu_xqnode & tNode, Setup_t & tSetup )
{
	// CreateMultiNode (plain words branch), src/searchnode.cpp:986-1042 + ExtNWay_T ctor/ConstructNode :3767-3802
	std::vector<TermNode_c*> dNodes;
	for ( int i=0; i<tNode.n_words; ++i )
	{
		TermNode_c * p = CreateTerm ( tNode, tNode.first_word+i, tSetup );
		if ( !p )
		{
			for ( auto q : dNodes ) delete q;
			return nullptr;
		}
		dNodes.push_back ( p );
	}
	std::vector<int> dAtomPos;
	for ( auto p : dNodes )
		dAtomPos.push_back ( p->m_iAtomPos );

	std::vector<WORD> dPositions ( dNodes.size() );
	for ( size_t i=0; i<dPositions.size(); ++i )
		dPositions[i] = (WORD)i;
	RefSort ( dPositions, [&] ( WORD a, WORD b ) { return dNodes[a]->GetDocsCount()<dNodes[b]->GetDocsCount(); } );

	WORD uLPos = dPositions[0];
	Node_c * pCur = dNodes[uLPos++];
	AndNode_c * pCurEx = nullptr;
	for ( size_t i=1; i<dNodes.size(); ++i )
	{
		WORD uRPos = dPositions[i];
		pCurEx = new AndNode_c;
		pCurEx->m_pLeft.reset ( pCur );
		pCurEx->m_pRight.reset ( dNodes[uRPos++] );
		pCurEx->m_uNodePosL = uLPos;
		pCurEx->m_uNodePosR = uRPos;
		uLPos = 0;
		pCur = pCurEx;
	}
	if ( pCurEx )
		pCurEx->m_bQPosReverse = true;

	auto * pRes = new NWayNode_c<FSM>;
	pRes->m_pNode.reset ( pCur );
	pRes->m_iAtomPos = dNodes[0]->m_iAtomPos;
	pRes->Init ( dAtomPos, tNode.oparg );
	return pRes;
}

static Node_c * CreateNode ( int iNode, Setup_t & tSetup )
{
	const mgpu_query & q = *tSetup.m_pQuery;
	if ( iNode<0 || iNode>=q.n_nodes )
	{
		tSetup.m_iError = MGPU_E_BAD_QUERY;
		return nullptr;
	}
	const mgpu_xqnode & tNode = q.nodes[iNode];

	if ( tNode.n_words )
	{
		if ( tNode.n_words==1 )
			return CreateTerm ( tNode, tNode.first_word, tSetup );
		switch ( tNode.op )
		{
		case MGPU_OP_PHRASE:	return CreateMultiNode<FSMphrase_c> ( tNode, tSetup );
		case MGPU_OP_PROXIMITY:	return CreateMultiNode<FSMproximity_c> ( tNode, tSetup );
		case MGPU_OP_QUORUM:
			{
				// src/searchnode.cpp:1638-1688: threshold >= words (or > 256 words) -> AND chain, threshold 1 -> OR chain over the keywords
				// sorted by doc count; everything else (incl. a percent threshold that rounds to 0) is a real ExtQuorum_c
				const int iCount = tNode.n_words, iThr = tNode.oparg;
				const bool bOr = ( iThr<iCount && iCount<=256 && iThr==1 );
				if ( iThr<iCount && iCount<=256 && iThr!=1 )
				{
					std::vector<TermNode_c*> dWords;
					for ( int i=0; i<iCount; ++i )
						dWords.push_back ( CreateTerm ( tNode, tNode.first_word+i, tSetup ) );
					QuorumNode_c * pQuorum = new QuorumNode_c;
					pQuorum->Init ( dWords, iThr );
					return pQuorum;
				}
				std::vector<Node_c*> dTerms;
				for ( int i=0; i<iCount; ++i )
					dTerms.push_back ( CreateTerm ( tNode, tNode.first_word+i, tSetup ) );
				RefSort ( dTerms, [] ( Node_c * a, Node_c * b ) { return a->GetDocsCount()<b->GetDocsCount(); } );
				Node_c * pCur = dTerms[0];
				for ( size_t i=1; i<dTerms.size(); ++i )
				{
					TwoferNode_c * pTwo = bOr ? (TwoferNode_c*)new OrNode_c : (TwoferNode_c*)new AndNode_c;
					pTwo->m_pLeft.reset ( pCur );
					pTwo->m_pRight.reset ( dTerms[i] );
					pCur = pTwo;
				}
				return pCur;
			}
		default:				tSetup.m_iError = MGPU_E_UNSUPPORTED; return nullptr;
		}
	}

	const int nChildren = tNode.n_children;
	if ( nChildren<1 )
		return nullptr;	// empty node
	const int32_t * pChildren = q.children + tNode.first_child;

	// AND over single-word children -> ExtMultiAnd_T (:1711-1776)
	bool bAndTerms = ( tNode.op==MGPU_OP_AND );
	for ( int i=0; i<nChildren && bAndTerms; ++i )
		bAndTerms = ( q.nodes[pChildren[i]].n_words==1 );
	bool bMultiAnd = bAndTerms && nChildren>1;
	for ( int i=0; i<nChildren && bMultiAnd; ++i )
	{
		const mgpu_xqnode & tChild = q.nodes[pChildren[i]];
		const mgpu_xqkeyword & tWord = q.words[tChild.first_word];
		if ( tWord.field_start || tWord.field_end || tChild.field_max_pos )
			bMultiAnd = false;	// src/searchnode.cpp:1716-1727
	}
	if ( bAndTerms && nChildren>1 && !bMultiAnd )
	{
		// terms sorted by frequency, chained with ExtAnd_c (src/searchnode.cpp:1734-1762)
		std::vector<Node_c*> dTerms;
		for ( int i=0; i<nChildren; ++i )
			dTerms.push_back ( CreateNode ( pChildren[i], tSetup ) );
		RefSort ( dTerms, [] ( Node_c * a, Node_c * b ) { return a->GetDocsCount()<b->GetDocsCount(); } );
		Node_c * pCur = dTerms[0];
		for ( size_t i=1; i<dTerms.size(); ++i )
		{
			AndNode_c * pAnd = new AndNode_c;
			pAnd->m_pLeft.reset ( pCur );
			pAnd->m_pRight.reset ( dTerms[i] );
			pCur = pAnd;
		}
		return pCur;
	}
	if ( bMultiAnd )
	{
		auto * p = new MultiAndNode_c;
		p->m_bUseBM25 = tSetup.m_bUseBM25;
		for ( int i=0; i<nChildren; ++i )
		{
			const mgpu_xqnode & tChild = q.nodes[pChildren[i]];
			const mgpu_xqkeyword & tWord = q.words[tChild.first_word];
			if ( tWord.field_start || tWord.field_end || tChild.field_max_pos )
			{
				tSetup.m_iError = MGPU_E_UNSUPPORTED;
				delete p;
				return nullptr;
			}
			p->m_dNodes.emplace_back();
			auto & n = p->m_dNodes.back();
			if ( n.m_tQword.Setup ( tSetup.m_pIndex, tWord.word ) )
				p->m_iNodesSet++;
			if ( q.shard_of_global && q.word_docs && q.word_docs[tChild.first_word]>=0 )
				n.m_tQword.m_iOrderDocs = q.word_docs[tChild.first_word];
			n.m_tQword.m_sWord = tWord.word;
			n.m_tQword.m_pIndex = tSetup.m_pIndex;
			n.m_tQword.m_iAtomPos = tWord.atom_pos;
			n.m_tQword.m_fBoost = tWord.boost;
			n.m_tQword.m_bExcluded = tWord.excluded!=0;
			n.m_tQword.m_bExpanded = tWord.expanded!=0;
			n.m_iAtomPos = tWord.atom_pos;
			n.m_uNodepos = (WORD)i;
			n.m_bNotWeighted = tChild.not_weighted!=0;
			n.m_uQueriedFields = tChild.field_mask;
			n.m_fIDF = 0.0f;
			n.m_tRowID = INVALID_ROWID;
			n.m_iWordIdx = tChild.first_word;
		}
		p->Finalize();
		return p;
	}
	if ( bAndTerms )
		return CreateNode ( pChildren[0], tSetup );	// degenerate 1-child AND: generic create returns the child

	if ( tNode.op==MGPU_OP_SENTENCE || tNode.op==MGPU_OP_PARAGRAPH )
	{
		// generic create, src/searchnode.cpp:1785-1806: pCur = new ExtUnit_c ( pCur, pNext, field mask, boundary keyword )
		Node_c * pCur = nullptr;
		for ( int i=0; i<nChildren; ++i )
		{
			Node_c * pNext = CreateNode ( pChildren[i], tSetup );
			if ( tSetup.m_iError!=MGPU_OK )
			{
				delete pNext; delete pCur;
				return nullptr;
			}
			if ( !pNext ) continue;
			if ( !pCur ) { pCur = pNext; continue; }
			auto * pUnit = new UnitNode_c;
			pUnit->m_iAtomPos = pCur->m_iAtomPos;
			pUnit->m_pLeft.reset ( pCur );
			pUnit->m_pRight.reset ( pNext );
			TermNode_c * pDot = new TermNode_c;
			const char * sUnit = tNode.op==MGPU_OP_SENTENCE ? "\3sentence" : "\3paragraph";
			pDot->m_tQword.Setup ( tSetup.m_pIndex, sUnit );
			pDot->m_tQword.m_sWord = sUnit;
			pDot->m_tQword.m_pIndex = tSetup.m_pIndex;
			pDot->m_uQueriedFields = tNode.field_mask;
			pDot->m_bNotWeighted = true;
			pDot->m_bUseBM25 = tSetup.m_bUseBM25;
			pUnit->m_pDot.reset ( pDot );
			pCur = pUnit;
		}
		return pCur;
	}

	if ( tNode.op==MGPU_OP_NOTNEAR )
	{
		// generic create, src/searchnode.cpp:1785-1806: ExtNotNear_c ( must, not, distance )
		if ( nChildren!=2 )
		{
			tSetup.m_iError = MGPU_E_BAD_QUERY;
			return nullptr;
		}
		Node_c * pMust = CreateNode ( pChildren[0], tSetup );
		Node_c * pNot = tSetup.m_iError==MGPU_OK ? CreateNode ( pChildren[1], tSetup ) : nullptr;
		if ( tSetup.m_iError!=MGPU_OK )
		{
			delete pMust;
			delete pNot;
			return nullptr;
		}
		if ( !pMust || !pNot )
			return pMust ? pMust : pNot;	// an empty child drops out of the fold (:1788-1794)
		auto * pRes = new NotNearNode_c;
		pRes->m_pLeft.reset ( pMust );
		pRes->m_pRight.reset ( pNot );
		pRes->m_iDist = tNode.oparg;
		pRes->m_iAtomPos = pMust->m_iAtomPos;
		return pRes;
	}

	if ( tNode.op==MGPU_OP_BEFORE )
	{
		// CreateOrderNode, src/searchnode.cpp:1044-1074
		if ( nChildren<2 )
			return nullptr;	// "order node requires at least two children"
		auto * pOrder = new OrderNode_c;
		for ( int i=0; i<nChildren; ++i )
		{
			Node_c * p = CreateNode ( pChildren[i], tSetup );
			if ( !p || tSetup.m_iError!=MGPU_OK )
			{
				delete pOrder;
				return nullptr;	// "failed to create order node, hitlist unavailable"
			}
			pOrder->m_dChildren.emplace_back ( p );
		}
		pOrder->m_dDoc.assign ( nChildren, ExtDoc_t { INVALID_ROWID, 0, 0.0f } );
		pOrder->m_dHas.assign ( nChildren, 0 );
		pOrder->m_dChildHits.resize ( nChildren );
		pOrder->m_iAtomPos = pOrder->m_dChildren[0]->m_iAtomPos;
		return pOrder;
	}

	if ( tNode.op==MGPU_OP_NEAR )
	{
		// CreateMultiNode<ExtMultinear_c> (children branch, src/searchnode.cpp:933-978) + ExtNWay_T ctor / ConstructNode (:3767-3802):
		// the children, sorted by doc count, are chained with ExtAnd_c that stamp each hit with its child's 1-based position
		std::vector<Node_c*> dNodes;
		for ( int i=0; i<nChildren; ++i )
		{
			Node_c * p = CreateNode ( pChildren[i], tSetup );
			if ( tSetup.m_iError!=MGPU_OK )
			{
				for ( auto q2 : dNodes ) delete q2;
				return nullptr;
			}
			if ( p )
				dNodes.push_back ( p );
		}
		if ( dNodes.size()<2 )
		{
			for ( auto q2 : dNodes ) delete q2;
			return nullptr;	// "can't create phrase node, hitlists unavailable"
		}
		std::vector<WORD> dPositions ( dNodes.size() );
		for ( size_t i=0; i<dPositions.size(); ++i )
			dPositions[i] = (WORD)i;
		RefSort ( dPositions, [&] ( WORD a, WORD b ) { return dNodes[a]->GetDocsCount()<dNodes[b]->GetDocsCount(); } );
		WORD uLPos = dPositions[0];
		Node_c * pChain = dNodes[uLPos++];
		AndNode_c * pLastAnd = nullptr;
		for ( size_t i=1; i<dNodes.size(); ++i )
		{
			WORD uRPos = dPositions[i];
			pLastAnd = new AndNode_c;
			pLastAnd->m_pLeft.reset ( pChain );
			pLastAnd->m_pRight.reset ( dNodes[uRPos++] );
			pLastAnd->m_uNodePosL = uLPos;
			pLastAnd->m_uNodePosR = uRPos;
			uLPos = 0;
			pChain = pLastAnd;
		}
		pLastAnd->m_bQPosReverse = true;
		auto * pNear = new NWayNode_c<FSMmultinear_c>;
		pNear->m_pNode.reset ( pChain );
		pNear->m_iAtomPos = dNodes[0]->m_iAtomPos;
		pNear->InitNear ( (int)dNodes.size(), tNode.oparg, false );
		return pNear;
	}

	if ( tNode.op==MGPU_OP_AND )
	{
		// AND over non-terms: children sorted by doc count, chain of ExtAnd_c (:1745-1770 applies to term-only ANDs that
		// cannot be multi-AND; the generic fold below (:1785-1806) is what mixed children get)
	}

	Node_c * pCur = nullptr;
	for ( int i=0; i<nChildren; ++i )
	{
		Node_c * pNext = CreateNode ( pChildren[i], tSetup );
		if ( tSetup.m_iError!=MGPU_OK )
		{
			delete pNext; delete pCur;
			return nullptr;
		}
		if ( !pNext ) continue;
		if ( !pCur ) { pCur = pNext; continue; }
		TwoferNode_c * pTwo = nullptr;
		switch ( tNode.op )
		{
		case MGPU_OP_OR:		pTwo = new OrNode_c; break;
		case MGPU_OP_MAYBE:		pTwo = new MaybeNode_c; break;
		case MGPU_OP_AND:		pTwo = new AndNode_c; break;
		case MGPU_OP_ANDNOT:	pTwo = new AndNotNode_c; break;
		default:
			tSetup.m_iError = MGPU_E_UNSUPPORTED;
			delete pNext; delete pCur;
			return nullptr;
		}
		pTwo->m_pLeft.reset ( pCur );
		pTwo->m_pRight.reset ( pNext );
		pCur = pTwo;
	}
	return pCur;
}

//////////////////////////////////////////////////////////////////////////
// ranking + sorting
//////////////////////////////////////////////////////////////////////////

/// RankerState_Proximity_fn<true,HANDLE_DUPES>, src/sphinxsearch.cpp:1319-1438
struct ProximityState_t
{
	BYTE m_uLCS[256];
	BYTE m_uCurLCS = 0;
	int m_iExpDelta = -INT_MAX;
	int m_iLastHitPosWithField = -INT_MAX;
	int m_iFields = 0;
	const int * m_pWeights = nullptr;
	bool m_bDupes = false;
	DWORD m_uLcsTailPos = 0, m_uLcsTailQposMask = 0, m_uCurQposMask = 0, m_uCurPos = 0;

	void Init ( int iFields, const int * pWeights, bool bDupes )
	{
		memset ( m_uLCS, 0, sizeof(m_uLCS) );
		m_iFields = iFields;
		m_pWeights = pWeights;
		m_bDupes = bDupes;
	}
	void Update ( const ExtHit_t * pHlist )
	{
		if ( !m_bDupes )
		{
			const int iPosWithField = (int)HitPosWithField ( pHlist->m_uHitpos );
			int iDelta = iPosWithField - pHlist->m_uQuerypos;
			if ( iPosWithField>m_iLastHitPosWithField )
				m_uCurLCS = (BYTE)( ( ( iDelta==m_iExpDelta ) ? m_uCurLCS : 0 ) + BYTE(pHlist->m_uWeight) );
			DWORD uField = (DWORD)HitField ( pHlist->m_uHitpos );
			if ( m_uCurLCS>m_uLCS[uField] )
				m_uLCS[uField] = m_uCurLCS;
			m_iLastHitPosWithField = iPosWithField;
			m_iExpDelta = iDelta + pHlist->m_uSpanlen - 1;
		} else
		{
			DWORD uPos = HitPosWithField ( pHlist->m_uHitpos );
			DWORD uField = (DWORD)HitField ( pHlist->m_uHitpos );
			if ( (DWORD)HitField ( m_uCurPos )!=uField )
				m_uCurQposMask = 0;
			if ( uPos!=m_uCurPos )
			{
				if ( m_uCurLCS<2 )
				{
					m_uLcsTailPos = m_uCurPos;
					m_uLcsTailQposMask = m_uCurQposMask;
					m_uCurLCS = 1;
				}
				m_uCurQposMask = 0;
				m_uCurPos = uPos;
				if ( m_uLCS[uField]<pHlist->m_uWeight )
					m_uLCS[uField] = BYTE(pHlist->m_uWeight);
			}
			m_uCurQposMask |= ( 1UL<<pHlist->m_uQuerypos );
			int iDelta = (int)( m_uCurPos-m_uLcsTailPos );
			if ( iDelta && iDelta<32 && ( m_uCurQposMask>>iDelta ) & m_uLcsTailQposMask )
			{
				m_uLcsTailQposMask = ( 1UL<<pHlist->m_uQuerypos );
				m_uLcsTailPos = m_uCurPos;
				m_uCurLCS = BYTE ( m_uCurLCS+pHlist->m_uWeight );
				m_uCurQposMask = 0;
				if ( m_uCurLCS>m_uLCS[uField] )
					m_uLCS[uField] = m_uCurLCS;
			}
		}
	}
	int Finalize ( int iSeedWeight )
	{
		m_uCurLCS = 0;
		m_iExpDelta = -1;
		m_iLastHitPosWithField = -1;
		if ( m_bDupes )
		{
			m_uLcsTailPos = 0; m_uLcsTailQposMask = 0; m_uCurQposMask = 0; m_uCurPos = 0;
		}
		int iRank = 0;
		for ( int i=0; i<m_iFields; i++ )
		{
			iRank += (int)( m_uLCS[i] )*m_pWeights[i];
			m_uLCS[i] = 0;
		}
		return iSeedWeight + iRank*SPH_BM25_SCALE;
	}
};

/// RankerState_ProximityBM25Exact_fn (SPH04), src/sphinxsearch.cpp:1443-1536. NB m_uMinExpPos survives Finalize, like in the reference
struct Sph04State_t
{
	BYTE m_uLCS[256];
	BYTE m_uCurLCS = 0;
	int m_iExpDelta = -INT_MAX;
	int m_iLastHitPos = -1;
	DWORD m_uMinExpPos = 0;
	int m_iFields = 0;
	const int * m_pWeights = nullptr;
	DWORD m_uHeadHit = 0, m_uExactHit = 0;
	int m_iMaxQuerypos = 0;

	void Init ( int iFields, const int * pWeights, int iMaxQpos )
	{
		memset ( m_uLCS, 0, sizeof(m_uLCS) );
		m_iFields = iFields;
		m_pWeights = pWeights;
		m_iMaxQuerypos = iMaxQpos;
	}
	void Update ( const ExtHit_t * pHlist )
	{
		DWORD uField = (DWORD)HitField ( pHlist->m_uHitpos );
		int iPosWithField = (int)HitPosWithField ( pHlist->m_uHitpos );
		int iDelta = iPosWithField - pHlist->m_uQuerypos;
		const DWORD uPos = pHlist->m_uHitpos & 0x7FFFFFu;
		const bool bEnd = ( pHlist->m_uHitpos>>23 ) & 1u;
		if ( iDelta==m_iExpDelta && HitPosWithField ( pHlist->m_uHitpos )>=m_uMinExpPos )
		{
			if ( iPosWithField>m_iLastHitPos )
				m_uCurLCS = (BYTE)( m_uCurLCS + pHlist->m_uWeight );
			if ( bEnd && (int)pHlist->m_uQuerypos==m_iMaxQuerypos && (int)uPos==m_iMaxQuerypos )
				m_uExactHit |= ( 1UL<<uField );
		} else
		{
			if ( iPosWithField>m_iLastHitPos )
				m_uCurLCS = BYTE(pHlist->m_uWeight);
			if ( uPos==1 )
			{
				m_uHeadHit |= ( 1UL<<uField );
				if ( bEnd && m_iMaxQuerypos==1 )
					m_uExactHit |= ( 1UL<<uField );
			}
		}
		if ( m_uCurLCS>m_uLCS[uField] )
			m_uLCS[uField] = m_uCurLCS;
		m_iExpDelta = iDelta + pHlist->m_uSpanlen - 1;
		m_iLastHitPos = iPosWithField;
		m_uMinExpPos = HitPosWithField ( pHlist->m_uHitpos ) + 1;
	}
	int Finalize ( int iSeedWeight )
	{
		m_uCurLCS = 0;
		m_iExpDelta = -1;
		m_iLastHitPos = -1;
		int iRank = 0;
		for ( int i=0; i<m_iFields; i++ )
		{
			iRank += (int)( 4*m_uLCS[i] + 2*( ( m_uHeadHit>>i ) & 1 ) + ( ( m_uExactHit>>i ) & 1 ) )*m_pWeights[i];
			m_uLCS[i] = 0;
		}
		m_uHeadHit = 0;
		m_uExactHit = 0;
		return iSeedWeight + iRank*SPH_BM25_SCALE;
	}
};

struct Match_t
{
	RowID_t m_tRowID;
	int m_iWeight;
	int64_t m_dKeys[5];
};

struct Comparator_t
{
	int m_nKeys = 0;
	int m_dKind[5];
	int m_dDesc[5];
	// SPH_TEST_KEYPART / MatchRelevanceLt_fn, src/sphinxsort.cpp:4534-4790. returns "a is worse than b"
	bool IsLess ( const Match_t & a, const Match_t & b ) const
	{
		for ( int i=0; i<m_nKeys; ++i )
		{
			int64_t aa, bb;
			switch ( m_dKind[i] )
			{
			case MGPU_KEYPART_ROWID:	aa = a.m_tRowID; bb = b.m_tRowID; break;
			case MGPU_KEYPART_WEIGHT:	aa = a.m_iWeight; bb = b.m_iWeight; break;
			default:					aa = a.m_dKeys[i]; bb = b.m_dKeys[i]; break;
			}
			if ( aa!=bb )
				return ( m_dDesc[i]!=0 ) ^ ( aa>bb );
		}
		return a.m_tRowID>b.m_tRowID;
	}
};

/// CSphMatchQueue, src/sphinxsort.cpp:582-812: binary heap with the worst match at the root
struct MatchQueue_c
{
	std::vector<Match_t> m_dData;
	int m_iSize;
	int64_t m_iTotal = 0;
	const Comparator_t & m_tComp;

	MatchQueue_c ( int iSize, const Comparator_t & tComp ) : m_iSize ( iSize ), m_tComp ( tComp ) { m_dData.reserve ( iSize ); }
	void Push ( const Match_t & tEntry )
	{
		++m_iTotal;
		if ( (int)m_dData.size()==m_iSize )
		{
			if ( m_tComp.IsLess ( tEntry, m_dData[0] ) )
				return;
			Pop();
		}
		m_dData.push_back ( tEntry );
		int iEntry = (int)m_dData.size()-1;
		while ( iEntry )
		{
			int iParent = ( iEntry-1 )/2;
			if ( !m_tComp.IsLess ( m_dData[iEntry], m_dData[iParent] ) )
				break;
			std::swap ( m_dData[iEntry], m_dData[iParent] );
			iEntry = iParent;
		}
	}
	void Pop()
	{
		m_dData[0] = m_dData.back();
		m_dData.pop_back();
		int iEntry = 0, iUsed = (int)m_dData.size();
		while ( true )
		{
			int iChild = iEntry*2+1;
			if ( iChild>=iUsed )
				break;
			if ( iChild+1<iUsed && m_tComp.IsLess ( m_dData[iChild+1], m_dData[iChild] ) )
				++iChild;
			if ( m_tComp.IsLess ( m_dData[iChild], m_dData[iEntry] ) )
			{
				std::swap ( m_dData[iChild], m_dData[iEntry] );
				iEntry = iChild;
				continue;
			}
			break;
		}
	}
	/// best first
	void Flatten ( std::vector<Match_t> & dOut )
	{
		dOut.resize ( m_dData.size() );
		for ( int i=(int)dOut.size()-1; i>=0; --i )
		{
			dOut[i] = m_dData[0];
			Pop();
		}
	}
};

static bool HasQwordDupes ( const mgpu_query & q )
{
	// HasQwordDupes, src/sphinxsearch.cpp:4148-4164
	std::unordered_map<std::string,int> h;
	for ( int i=0; i<q.n_words; ++i )
		if ( !h.emplace ( q.words[i].word, 1 ).second )
			return true;
	return false;
}

static int SearchOne ( const Index_t & tIndex, const mgpu_query & q, mgpu_result & tRes )
{
	tRes.n_matches = 0;
	tRes.total_found = 0;
	if ( q.n_nodes<=0 || q.root<0 )
		return MGPU_OK;		// empty query matches nothing

	const bool bHitRanker = ( q.ranker==MGPU_RANK_PROXIMITY_BM25 || q.ranker==MGPU_RANK_WORDCOUNT );
	const mgpu_xqnode & tRoot = q.nodes[q.root];
	const bool bSingleWord = ( tRoot.n_words==1 && tRoot.n_children==0 );	// XQQuery_t::m_bSingleWord
	// sphCreateRanker, src/sphinxsearch.cpp:4189-4232
	const bool bStateRanker = q.ranker==MGPU_RANK_WORDCOUNT || q.ranker==MGPU_RANK_MATCHANY || q.ranker==MGPU_RANK_FIELDMASK || q.ranker==MGPU_RANK_SPH04
		|| ( ( q.ranker==MGPU_RANK_PROXIMITY_BM25 || q.ranker==MGPU_RANK_PROXIMITY ) && !bSingleWord );
	(void)bHitRanker;
	if ( q.ranker<MGPU_RANK_PROXIMITY_BM25 || q.ranker>MGPU_RANK_SPH04 )
		return MGPU_E_UNSUPPORTED;

	Setup_t tSetup { &tIndex, &q, q.ranker==MGPU_RANK_PROXIMITY_BM25 || q.ranker==MGPU_RANK_BM25 || q.ranker==MGPU_RANK_SPH04 };
	std::unique_ptr<Node_c> pRoot ( CreateNode ( q.root, tSetup ) );
	if ( tSetup.m_iError!=MGPU_OK )
		return tSetup.m_iError;

	// word stats for every query word, found or not
	if ( tRes.word_stats )
		for ( int i=0; i<q.n_words; ++i )
		{
			const WordEntry_t * pEntry = tIndex.FindWord ( q.words[i].word );
			tRes.word_stats[i].docs = pEntry ? pEntry->m_iDocs : 0;
			tRes.word_stats[i].hits = pEntry ? pEntry->m_iHits : 0;
		}
	if ( !pRoot )
		return MGPU_OK;

	// IDFs: sphCreateRanker, src/sphinxsearch.cpp:4293-4378
	QwordsHash_t hQwords;
	const int iMaxQpos = pRoot->GetQwords ( hQwords );
	const int iQwords = (int)hQwords.size();
	int64_t iTotalDocuments = q.total_docs>0 ? q.total_docs : tIndex.m_iTotalDocs;
	for ( auto & kv : hQwords )
	{
		ExtQword_t & tWord = kv.second;
		int64_t iTermDocs = tWord.m_iDocs;
		if ( q.word_docs && q.word_docs[tWord.m_iFirstWordIdx]>=0 )
			iTermDocs = q.word_docs[tWord.m_iFirstWordIdx];
		float fIDF = 0.0f;
		if ( iTermDocs )
		{
			const int64_t iTotalClamped = std::max ( iTotalDocuments, iTermDocs );
			float fLogTotal = logf ( float ( 1+iTotalClamped ) );
			if ( !q.plain_idf )
				fIDF = logf ( float ( iTotalClamped-iTermDocs+1 ) / float ( iTermDocs ) ) / ( 2*fLogTotal );
			else
				fIDF = logf ( float ( iTotalClamped ) / float ( iTermDocs ) ) / ( 2*fLogTotal );
		}
		if ( !q.unnormalized_tfidf )
			fIDF /= iQwords;
		tWord.m_fIDF = fIDF * tWord.m_fBoost;
	}
	pRoot->SetQwordsIDF ( hQwords );

	// field weights: BindWeights, src/sphinx.cpp:13903-13943 (already bound by the caller)
	const int iFields = (int)tIndex.m_dFields.size();
	std::vector<int> dWeights ( std::max ( iFields, 1 ), 1 );
	for ( int i=0; i<iFields && i<q.n_field_weights; ++i )
		dWeights[i] = q.field_weights[i];

	// sorter
	Comparator_t tComp;
	if ( q.n_sort_keys<=0 )
	{
		tComp.m_nKeys = 1;
		tComp.m_dKind[0] = MGPU_KEYPART_WEIGHT;
		tComp.m_dDesc[0] = 1;
	} else
	{
		if ( q.n_sort_keys>5 )
			return MGPU_E_BAD_QUERY;
		tComp.m_nKeys = q.n_sort_keys;
		for ( int i=0; i<q.n_sort_keys; ++i )
		{
			tComp.m_dKind[i] = q.sort_keys[i].kind;
			tComp.m_dDesc[i] = q.sort_keys[i].desc;
			if ( ( q.sort_keys[i].kind==MGPU_KEYPART_INT || q.sort_keys[i].kind==MGPU_KEYPART_FLOAT ) && ( q.sort_keys[i].attr<0 || q.sort_keys[i].attr>=(int)tIndex.m_dAttrs.size() ) )
				return MGPU_E_BAD_QUERY;
		}
	}
	const int iMaxMatches = q.max_matches>0 ? q.max_matches : 1000;
	MatchQueue_c tQueue ( iMaxMatches, tComp );
	const int iIndexWeight = q.index_weight ? q.index_weight : 1;

	ProximityState_t tProx;
	tProx.Init ( iFields, dWeights.data(), HasQwordDupes ( q ) && q.ranker!=MGPU_RANK_MATCHANY );	// MatchAny derives from <false,false>
	Sph04State_t tSph04;
	tSph04.Init ( iFields, dWeights.data(), iMaxQpos );
	// RankerState_MatchAny_fn, src/sphinxsearch.cpp:1582-1622
	int iPhraseK = 0;
	for ( int i=0; i<iFields; i++ )
		iPhraseK += dWeights[i]*iQwords;
	BYTE dMatchMask[256];
	memset ( dMatchMask, 0, sizeof(dMatchMask) );
	std::vector<ExtHit_t> dHits;
	const int iWeights = std::min ( iFields, 32 );

	ExtDoc_t tDoc;
	while ( pRoot->Next ( tDoc ) )
	{
		// EarlyReject + filters, src/sphinx.cpp:11903-11917
		bool bReject = false;
		for ( int f=0; f<q.n_filters && !bReject; ++f )
		{
			const mgpu_filter & tF = q.filters[f];
			int64_t v = tIndex.GetAttr ( tDoc.m_tRowID, tF.attr );
			bool bOk;
			if ( tF.kind==MGPU_FILTER_RANGE )
				bOk = ( v>=tF.min_value && v<=tF.max_value );
			else
			{
				bOk = false;
				for ( int k=0; k<tF.n_values && !bOk; ++k )
					bOk = ( tF.values[k]==v );
			}
			if ( tF.exclude )
				bOk = !bOk;
			bReject = !bOk;
		}
		if ( bReject )
			continue;

		int iWeight = 0;
		if ( tSetup.m_bUseBM25 )
			iWeight = (int)( ( tDoc.m_fTFIDF+0.5f )*SPH_BM25_SCALE );	// src/sphinxsearch.cpp:1070

		if ( bStateRanker )
		{
			dHits.clear();
			pRoot->CollectHits ( dHits );
			if ( dHits.empty() )
				continue;	// ExtRanker_State_T skips docs without hits, :1299-1304
			if ( q.ranker==MGPU_RANK_WORDCOUNT )
			{
				// RankerState_Wordcount_fn, src/sphinxsearch.cpp:1620-1643
				int iRank = 0;
				for ( const ExtHit_t & h : dHits )
					iRank += dWeights [ HitField ( h.m_uHitpos ) ];
				iWeight = iRank;
			} else if ( q.ranker==MGPU_RANK_FIELDMASK )
			{
				// RankerState_Fieldmask_fn, src/sphinxsearch.cpp:1648-1668
				DWORD uRank = 0;
				for ( const ExtHit_t & h : dHits )
					uRank |= 1UL<<HitField ( h.m_uHitpos );
				iWeight = (int)uRank;
			} else if ( q.ranker==MGPU_RANK_SPH04 )
			{
				for ( const ExtHit_t & h : dHits )
					tSph04.Update ( &h );
				iWeight = tSph04.Finalize ( iWeight );
			} else if ( q.ranker==MGPU_RANK_MATCHANY )
			{
				for ( const ExtHit_t & h : dHits )
				{
					tProx.Update ( &h );
					dMatchMask [ HitField ( h.m_uHitpos ) ] |= (BYTE)( 1<<( h.m_uQuerypos-1 ) );
				}
				// Finalize: the LCS state is reset by ProximityState_t::Finalize, so take the ranks first
				int iRank = 0;
				for ( int i=0; i<iFields; i++ )
				{
					if ( dMatchMask[i] )
						iRank += (int)( __builtin_popcount ( dMatchMask[i] ) + ( tProx.m_uLCS[i]-1 )*iPhraseK )*dWeights[i];
					dMatchMask[i] = 0;
				}
				tProx.Finalize ( 0 );
				iWeight = iRank;
			} else
			{
				for ( const ExtHit_t & h : dHits )
					tProx.Update ( &h );
				iWeight = tProx.Finalize ( iWeight );
				if ( q.ranker==MGPU_RANK_PROXIMITY )
					iWeight /= SPH_BM25_SCALE;	// RankerState_Proximity_fn<false,..>::Finalize returns the bare rank (seed is 0 without BM25)
			}
		} else if ( q.ranker==MGPU_RANK_NONE )
			iWeight = 1;
		else if ( q.ranker==MGPU_RANK_PROXIMITY )
		{
			// single keyword: ExtRanker_WeightSum_c<false>, src/sphinxsearch.cpp:1131-1134
			DWORD uRank = 0;
			DWORD uMask = tDoc.m_uDocFields;
			if ( !uMask )
				uRank = 1;
			else
				for ( int i=0; i<iWeights; i++ )
					if ( uMask & ( 1u<<i ) )
						uRank += (DWORD)dWeights[i];
			iWeight = (int)uRank;
		} else
		{
			// ExtRanker_WeightSum_c, :1096-1141
			DWORD uRank = 0;
			DWORD uMask = tDoc.m_uDocFields;
			if ( !uMask )
				uRank = 1;
			else
				for ( int i=0; i<iWeights; i++ )
					if ( uMask & ( 1u<<i ) )
						uRank += (DWORD)dWeights[i];
			iWeight = (int)( (DWORD)iWeight + uRank*SPH_BM25_SCALE );
		}

		// MatchExtended, src/sphinx.cpp:12190-12269
		if ( tIndex.IsDead ( tDoc.m_tRowID ) )
			continue;
		Match_t tMatch;
		tMatch.m_tRowID = tDoc.m_tRowID;
		tMatch.m_iWeight = iWeight*iIndexWeight;
		for ( int i=0; i<tComp.m_nKeys; ++i )
		{
			tMatch.m_dKeys[i] = ( tComp.m_dKind[i]==MGPU_KEYPART_INT || tComp.m_dKind[i]==MGPU_KEYPART_FLOAT ) ? tIndex.GetAttr ( tDoc.m_tRowID, q.sort_keys[i].attr ) : 0;
			if ( tComp.m_dKind[i]==MGPU_KEYPART_FLOAT )
			{
				// SPH_KEYPART_FLOAT, src/sphinxsort.cpp:4690-4696: compared as floats. Kept as an integer of the same order so that the
				// comparator stays a strict weak order when a row holds a NaN (the reference's aa>bb is false both ways there)
				DWORD u = (DWORD)tMatch.m_dKeys[i];
				float f; memcpy ( &f, &u, 4 );
				if ( f==0.0f )
					u = 0;
				tMatch.m_dKeys[i] = ( u & 0x80000000u ) ? (int64_t)(DWORD)~u : (int64_t)( u | 0x80000000u );
			}
		}
		tQueue.Push ( tMatch );
	}

	std::vector<Match_t> dOut;
	tQueue.Flatten ( dOut );
	tRes.n_matches = (int)dOut.size();
	tRes.total_found = tQueue.m_iTotal;
	int iFirstIntKey = -1;
	for ( int i=0; i<tComp.m_nKeys && iFirstIntKey<0; ++i )
		if ( tComp.m_dKind[i]==MGPU_KEYPART_INT )
			iFirstIntKey = i;
	for ( size_t i=0; i<dOut.size(); ++i )
	{
		tRes.rowid[i] = dOut[i].m_tRowID;
		tRes.weight[i] = dOut[i].m_iWeight;
		if ( tRes.docid )
			tRes.docid[i] = tIndex.GetAttr ( dOut[i].m_tRowID, 0 );
		if ( tRes.sort_attr )
			tRes.sort_attr[i] = iFirstIntKey>=0 ? dOut[i].m_dKeys[iFirstIntKey] : 0;
	}
	return MGPU_OK;
}

} // namespace

//////////////////////////////////////////////////////////////////////////
// C entry points (ctypes)
//////////////////////////////////////////////////////////////////////////

extern "C"
{

struct oracle_index { Index_t m_t; };

oracle_index * oracle_open ( const char * szPrefix, char * szErr, int iErrLen )
{
	auto * p = new oracle_index;
	if ( !p->m_t.Open ( szPrefix ) )
	{
		if ( szErr && iErrLen>0 )
			snprintf ( szErr, iErrLen, "%s", p->m_t.m_sError.c_str() );
		delete p;
		return nullptr;
	}
	return p;
}

void oracle_close ( oracle_index * p )
{
	delete p;
}

int64_t oracle_total_docs ( const oracle_index * p )		{ return p->m_t.m_iTotalDocs; }
int oracle_num_fields ( const oracle_index * p )			{ return (int)p->m_t.m_dFields.size(); }

/// same query/result structs as the product's C ABI (include/mgpu.h); processes queries [0,n) one by one
int oracle_search_batch ( oracle_index * p, const mgpu_query * pQueries, int nQueries, mgpu_result * pResults )
{
	for ( int i=0; i<nQueries; ++i )
		pResults[i].status = SearchOne ( p->m_t, pQueries[i], pResults[i] );
	return MGPU_OK;
}

/// DiskIndexQword_c::ReadNext over a whole doclist -- the checker for the GPU block decoder (kernel K1)
int oracle_decode_doclist ( oracle_index * p, const char * szWord, uint32_t * pRowid, uint32_t * pHits, uint32_t * pFields, uint64_t * pHitlistPos, int64_t iCapacity, int64_t * pOut )
{
	Qword_t q;
	*pOut = 0;
	if ( !q.Setup ( &p->m_t, szWord ) )
		return 0;
	int64_t n = 0;
	while ( true )
	{
		q.ReadNext();
		if ( q.m_tRowID==INVALID_ROWID )
			break;
		if ( n<iCapacity )
		{
			pRowid[n] = q.m_tRowID;
			pHits[n] = q.m_uMatchHits;
			pFields[n] = q.m_uFields;
			pHitlistPos[n] = q.m_iHitlistPos;
		}
		++n;
	}
	*pOut = n;
	return 1;
}

/// hit positions of one (word, doc): GetNextHit loop -- checker for the GPU hit decoder
int oracle_decode_hitlist ( oracle_index * p, const char * szWord, uint64_t uHitlistPos, uint32_t * pHits, int iCapacity )
{
	Qword_t q;
	if ( !q.Setup ( &p->m_t, szWord ) )
		return -1;
	q.SeekHitlist ( uHitlistPos );
	int n = 0;
	while ( true )
	{
		Hitpos_t u = q.GetNextHit();
		if ( u==EMPTY_HIT )
			break;
		if ( n<iCapacity )
			pHits[n] = u;
		++n;
	}
	return n;
}

int oracle_word_stats ( oracle_index * p, const char * szWord, int64_t * pDocs, int64_t * pHits )
{
	const WordEntry_t * pEntry = p->m_t.FindWord ( szWord );
	if ( !pEntry )
		return 0;
	*pDocs = pEntry->m_iDocs;
	*pHits = pEntry->m_iHits;
	return 1;
}

} // extern "C"
