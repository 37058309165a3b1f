// Index format v62 structures shared by the writer and the loader.
// Reference: header writer src/sphinx.cpp:8781-8891 (IndexWriteHeader), reader :13207-13392 (LoadHeader);
// settings blocks src/indexsettings.cpp:303-333 (tokenizer), :405-453 (dict), :506-523 (field filter).
#pragma once
#include <string.h>

#include "vbyte.h"
#include <string>
#include <vector>

namespace mgpu
{

static const DWORD INDEX_MAGIC_HEADER = 0x58485053;	// 'SPHX', src/sphinxint.h:45
static const DWORD INDEX_FORMAT_VERSION = 62;			// src/sphinxint.h:46
static const int SPH_WORDLIST_CHECKPOINT = 64;			// src/sphinxint.h:1663
static const int DOCLIST_HINT_THRESH = 256;			// src/indexformat.h:20
static const int DOCINFO_INDEX_FREQ = 128;				// src/sphinxint.h:312

enum { SPH_ATTR_INTEGER = 1, SPH_ATTR_TIMESTAMP = 2, SPH_ATTR_BIGINT = 6 };	// src/sphinxexpr.h:31-37
enum { SPH_HIT_FORMAT_PLAIN = 0, SPH_HIT_FORMAT_INLINE = 1 };					// src/indexsettings.h:222
enum { SPH_HITLESS_NONE = 0, SPH_HITLESS_SOME = 1, SPH_HITLESS_ALL = 2 };		// src/indexsettings.h:214
enum { TOKENIZER_UTF8 = 2, TOKENIZER_NGRAM = 3 };								// src/indexsettings.h:32
enum { FIELD_INDEXED = 2 };														// src/sphinx.h:1324

struct SchemaField_t
{
	std::string	m_sName;
	DWORD		m_uFlags = FIELD_INDEXED;
	BYTE		m_bPayload = 0;
};

struct SchemaAttr_t
{
	std::string	m_sName;
	DWORD		m_eType = SPH_ATTR_INTEGER;
	DWORD		m_iBitOffset = 0;	// CSphAttrLocator (src/sphinx.h:993-1014)
	DWORD		m_iBitCount = 32;
	BYTE		m_bPayload = 0;
	DWORD		m_uFlags = 0;
};

/// everything IndexWriteHeader stores that the query path needs (plus what it must skip over)
struct IndexHeader_t
{
	DWORD						m_uVersion = INDEX_FORMAT_VERSION;
	std::vector<SchemaField_t>	m_dFields;
	std::vector<SchemaAttr_t>	m_dAttrs;

	int64_t		m_iDictCheckpointsOffset = 0;
	DWORD		m_iDictCheckpoints = 0;
	BYTE		m_iInfixCodepointBytes = 0;
	DWORD		m_iInfixBlocksOffset = 0;
	DWORD		m_iInfixBlocksWordsSize = 0;

	DWORD		m_iTotalDocuments = 0;
	int64_t		m_iTotalBytes = 0;

	// CSphIndexSettings subset
	DWORD		m_iMinPrefixLen = 0, m_iMinInfixLen = 0, m_iMaxSubstringLen = 0;
	DWORD		m_eHitless = SPH_HITLESS_NONE;
	DWORD		m_eHitFormat = SPH_HIT_FORMAT_INLINE;
	DWORD		m_iBoundaryStep = 0, m_iStopwordStep = 1, m_iOvershortStep = 1, m_iEmbeddedLimit = 16384;
	BYTE		m_bIndexFieldLens = 0;
	DWORD		m_iSkiplistBlockSize = 32;

	// tokenizer / dict settings subset
	DWORD		m_iMinWordLen = 1;
	BYTE		m_bWordDict = 1;

	int64_t		m_iDocinfo = 0;			// rows in .spa
	int64_t		m_iDocinfoIndex = 0;	// min-max blocks
	int64_t		m_iMinMaxIndex = 0;		// DWORD offset of the min-max rows inside .spa

	int			RowStride() const		// in DWORDs
	{
		DWORD uBits = 0;
		for ( const auto & a : m_dAttrs )
			if ( a.m_iBitOffset+a.m_iBitCount>uBits )
				uBits = a.m_iBitOffset+a.m_iBitCount;
		return (int)( ( uBits+31 )/32 );
	}
};

/// CSphDictEntry (src/sphinx.h) as stored in .spi
/// sphFNV64 over a zero-terminated keyword (src/fnv64.cpp:16-50): the word id of dict=crc indexes (CCRCEngine<false>::DoCrc, src/sphinx.cpp:17318)
inline uint64_t WordIdFNV64 ( const char * sWord )
{
	uint64_t h = 0xcbf29ce484222325ull;
	for ( const unsigned char * p = (const unsigned char *)sWord; *p; ++p )
	{
		h ^= (uint64_t)*p;
		h *= 0x100000001b3ull;
	}
	return h;
}

/// how a dict=crc entry is named in the loaders' keyword maps: a zero byte (no keyword starts with one) + the 8 bytes of its word id
inline std::string CrcDictKey ( uint64_t uWordID )
{
	std::string s ( 9, '\0' );
	memcpy ( &s[1], &uWordID, 8 );
	return s;
}

struct DictEntry_t
{
	std::string	m_sKeyword;				// dict=crc: CrcDictKey ( m_uWordID )
	uint64_t	m_uWordID = 0;			// dict=crc only
	int64_t		m_iDoclistOffset = 0;
	int64_t		m_iDoclistLength = 0;	// bytes incl. the terminating zero varint
	int			m_iDocs = 0;
	int			m_iHits = 0;
	int64_t		m_iSkiplistOffset = 0;	// valid if m_iDocs > skiplist block
	int64_t		m_iSkiplistLength = 0;	// bytes in .spe (derived; not stored in the dict)
};

void	WriteHeader ( ByteBuf_t & tOut, const IndexHeader_t & tHdr );
bool	ReadHeader ( const BYTE * pData, size_t iLen, IndexHeader_t & tHdr, std::string & sError );

/// sphDoclistHintPack (src/sphinx.cpp:10864-10878)
BYTE	DoclistHintPack ( int64_t iDocs, int64_t iLen );

/// keywords-dictionary (.spi, dict=keywords) writer: CSphDictKeywords::DictEnd (src/sphinx.cpp:19462-19587)
/// with CSphKeywordDeltaWriter (src/sphinxint.h:1606-1658). Entries must arrive sorted by strcmp.
struct DictWriter_c
{
	ByteBuf_t	m_tOut;
	int			m_iSkiplistBlockSize = 32;
	int			m_iWords = 0;
	std::string	m_sLast;
	struct Checkpoint_t { std::string m_sWord; int64_t m_iOffset; };
	std::vector<Checkpoint_t> m_dCheckpoints;

	/// bCrc: the dict=crc form (CSphDiskDictTraits::DictEntry / DictEndEntries / DictEnd, src/sphinx.cpp:18263-18339): entries sorted by
	/// word id, delta-coded ids and doclist offsets, checkpoints of {word id, offset}; Finish needs the end of the last doclist then
	explicit DictWriter_c ( int iSkiplistBlockSize, bool bCrc=false );
	void	AddEntry ( const DictEntry_t & tEntry );
	void	Finish ( IndexHeader_t & tHdr, int64_t iDoclistEnd=0 );
	bool		m_bCrc = false;
	uint64_t	m_uLastWordID = 0;
	int64_t		m_iLastDoclistPos = 0;
	std::vector<std::pair<uint64_t,int64_t>> m_dCrcCheckpoints;
};

/// reads every dictionary entry of a dict=keywords or dict=crc .spi
/// (KeywordsBlockReader_c::UnpackWord src/indexformat.cpp:641-691; CWordlist::GetWord :425-473)
bool	ReadDictionary ( const BYTE * pSpi, size_t iLen, const IndexHeader_t & tHdr, std::vector<DictEntry_t> & dOut, std::string & sError );

} // namespace mgpu
