// Index format v62: .sph header and .spi dictionary (dict=keywords) codec.
// See index_format.h for the reference line citations of every block.
#include "index_format.h"

#include <algorithm>

namespace mgpu
{

static void WriteFileInfo ( ByteBuf_t & t )
{
	// WriteFileInfo (src/indexsettings.cpp:1184): size, ctime, mtime, crc32 of a file that does not exist
	t.PutOffset ( 0 ); t.PutOffset ( 0 ); t.PutOffset ( 0 ); t.PutDword ( 0 );
}

static void SkipFileInfo ( ByteReader_t & r )
{
	r.GetOffset(); r.GetOffset(); r.GetOffset(); r.GetDword();
}

void WriteHeader ( ByteBuf_t & t, const IndexHeader_t & h )
{
	// IndexWriteHeader, src/sphinx.cpp:8841-8891
	t.PutDword ( INDEX_MAGIC_HEADER );
	t.PutDword ( INDEX_FORMAT_VERSION );

	// WriteSchema, src/sphinx.cpp:8782-8812
	t.PutDword ( (DWORD)h.m_dFields.size() );
	for ( const auto & f : h.m_dFields )
	{
		t.PutString ( f.m_sName );
		t.PutDword ( f.m_uFlags );
		t.PutByte ( f.m_bPayload );
	}
	t.PutDword ( (DWORD)h.m_dAttrs.size() );
	for ( const auto & a : h.m_dAttrs )
	{
		t.PutString ( a.m_sName );
		t.PutDword ( a.m_eType );
		t.PutDword ( a.m_iBitOffset/32 );	// CalcRowitem(), kept for backwards compatibility
		t.PutDword ( a.m_iBitOffset );
		t.PutDword ( a.m_iBitCount );
		t.PutByte ( a.m_bPayload );
		t.PutDword ( a.m_uFlags );
	}

	// wordlist checkpoints
	t.PutOffset ( h.m_iDictCheckpointsOffset );
	t.PutDword ( h.m_iDictCheckpoints );
	t.PutByte ( h.m_iInfixCodepointBytes );
	t.PutDword ( h.m_iInfixBlocksOffset );
	t.PutDword ( h.m_iInfixBlocksWordsSize );

	// index stats
	t.PutDword ( h.m_iTotalDocuments );
	t.PutOffset ( h.m_iTotalBytes );

	// SaveIndexSettings, src/sphinx.cpp:8815-8841
	t.PutDword ( h.m_iMinPrefixLen );
	t.PutDword ( h.m_iMinInfixLen );
	t.PutDword ( h.m_iMaxSubstringLen );
	t.PutByte ( 0 );			// html_strip
	t.PutString ( "" );			// html_index_attrs
	t.PutString ( "" );			// html_remove_elements
	t.PutByte ( 0 );			// index_exact_words
	t.PutDword ( h.m_eHitless );
	t.PutDword ( h.m_eHitFormat );
	t.PutByte ( 0 );			// index_sp
	t.PutString ( "" );			// zones
	t.PutDword ( h.m_iBoundaryStep );
	t.PutDword ( h.m_iStopwordStep );
	t.PutDword ( h.m_iOvershortStep );
	t.PutDword ( h.m_iEmbeddedLimit );
	t.PutByte ( 0 );			// bigram_index
	t.PutString ( "" );			// bigram words
	t.PutByte ( h.m_bIndexFieldLens );
	t.PutByte ( 0 );			// preprocessor
	t.PutString ( "" );			// was: RLP context
	t.PutString ( "" );			// index token filter
	t.PutOffset ( 0 );			// blob update space
	t.PutDword ( h.m_iSkiplistBlockSize );
	t.PutString ( "" );			// hitless files

	// SaveTokenizerSettings, src/indexsettings.cpp:1197-1220
	t.PutByte ( TOKENIZER_UTF8 );
	t.PutString ( "" );			// charset_table (default)
	t.PutDword ( h.m_iMinWordLen );
	t.PutByte ( 1 );			// embedded synonyms (empty file <= embedded_limit)
	t.PutDword ( 0 );			// CSphTokenizerBase::WriteSynonyms: no exceptions
	t.PutString ( "" );			// synonyms file
	WriteFileInfo ( t );
	t.PutString ( "" );			// boundary
	t.PutString ( "" );			// ignore chars
	t.PutDword ( 0 );			// ngram len
	t.PutString ( "" );			// ngram chars
	t.PutString ( "" );			// blend chars
	t.PutString ( "" );			// blend mode

	// SaveDictionarySettings, src/indexsettings.cpp:1224-1272
	t.PutString ( "" );			// morphology
	t.PutString ( "" );			// morph fields
	t.PutByte ( 1 );			// embedded stopwords
	t.PutDword ( 0 );			// WriteStopwords: none
	t.PutString ( "" );			// stopwords
	t.PutDword ( 0 );			// stopword files
	t.PutByte ( 1 );			// embedded wordforms
	t.PutDword ( 0 );			// WriteWordforms: none
	t.PutDword ( 0 );			// wordform files
	t.PutDword ( 1 );			// min_stemming_len
	t.PutByte ( h.m_bWordDict );
	t.PutByte ( 0 );			// stopwords_unstemmed
	t.PutString ( "" );			// morph data fingerprint

	t.PutOffset ( h.m_iDocinfo );
	t.PutOffset ( h.m_iDocinfoIndex );
	t.PutOffset ( h.m_iMinMaxIndex );

	// CSphFieldFilterSettings::Save
	t.PutDword ( 0 );

	// average field lengths: only with index_field_lengths
	if ( h.m_bIndexFieldLens )
		for ( size_t i=0; i<h.m_dFields.size(); ++i )
			t.PutOffset ( 0 );
}


bool ReadHeader ( const BYTE * pData, size_t iLen, IndexHeader_t & h, std::string & sError )
{
	// CSphIndex_VLN::LoadHeader, src/sphinx.cpp:13252-13392
	ByteReader_t r ( pData, iLen );
	if ( r.GetDword()!=INDEX_MAGIC_HEADER )
	{
		sError = "not a SPHX index header";
		return false;
	}
	h.m_uVersion = r.GetDword();
	if ( h.m_uVersion<57 || h.m_uVersion>INDEX_FORMAT_VERSION )
	{
		sError = "unsupported index format version " + std::to_string ( h.m_uVersion ) + " (need 57..62)";
		return false;
	}

	// ReadSchema, src/sphinx.cpp:8722-8778
	DWORD nFields = r.GetDword();
	h.m_dFields.clear();
	for ( DWORD i=0; i<nFields && !r.m_bError; ++i )
	{
		SchemaField_t f;
		f.m_sName = r.GetString();
		f.m_uFlags = r.GetDword();
		f.m_bPayload = r.GetByte();
		h.m_dFields.push_back ( f );
	}
	DWORD nAttrs = r.GetDword();
	h.m_dAttrs.clear();
	for ( DWORD i=0; i<nAttrs && !r.m_bError; ++i )
	{
		SchemaAttr_t a;
		a.m_sName = r.GetString();
		a.m_eType = r.GetDword();
		r.GetDword();	// rowitem, ignored
		a.m_iBitOffset = r.GetDword();
		a.m_iBitCount = r.GetDword();
		a.m_bPayload = r.GetByte();
		if ( h.m_uVersion>=61 )
			a.m_uFlags = r.GetDword();
		h.m_dAttrs.push_back ( a );
	}

	h.m_iDictCheckpointsOffset = r.GetOffset();
	h.m_iDictCheckpoints = r.GetDword();
	h.m_iInfixCodepointBytes = r.GetByte();
	h.m_iInfixBlocksOffset = r.GetDword();
	h.m_iInfixBlocksWordsSize = r.GetDword();

	h.m_iTotalDocuments = r.GetDword();
	h.m_iTotalBytes = r.GetOffset();

	// LoadIndexSettings, src/sphinx.cpp:13207-13249
	h.m_iMinPrefixLen = r.GetDword();
	h.m_iMinInfixLen = r.GetDword();
	h.m_iMaxSubstringLen = r.GetDword();
	r.GetByte(); r.GetString(); r.GetString();
	r.GetByte();
	h.m_eHitless = r.GetDword();
	h.m_eHitFormat = r.GetDword();
	r.GetByte();
	r.GetString();
	h.m_iBoundaryStep = r.GetDword();
	h.m_iStopwordStep = r.GetDword();
	h.m_iOvershortStep = r.GetDword();
	h.m_iEmbeddedLimit = r.GetDword();
	r.GetByte(); r.GetString();
	h.m_bIndexFieldLens = r.GetByte();
	r.GetByte(); r.GetString();
	r.GetString();
	r.GetOffset();
	h.m_iSkiplistBlockSize = r.GetDword();
	if ( h.m_uVersion>=60 )
		r.GetString();

	// CSphTokenizerSettings::Load, src/indexsettings.cpp:303-333
	BYTE uTokType = r.GetByte();
	if ( uTokType!=TOKENIZER_UTF8 && uTokType!=TOKENIZER_NGRAM && !r.m_bError )
	{
		sError = "can't load an old index with SBCS tokenizer";
		return false;
	}
	r.GetString();
	h.m_iMinWordLen = r.GetDword();
	if ( r.GetByte() )
	{
		DWORD n = r.GetDword();
		for ( DWORD i=0; i<n && !r.m_bError; ++i )
			r.GetString();
	}
	r.GetString();
	SkipFileInfo ( r );
	r.GetString(); r.GetString(); r.GetDword(); r.GetString(); r.GetString(); r.GetString();

	// CSphDictSettings::Load, src/indexsettings.cpp:405-453
	r.GetString(); r.GetString();
	if ( r.GetByte() )
	{
		DWORD n = r.GetDword();
		for ( DWORD i=0; i<n && !r.m_bError; ++i )
			r.Unzip();
	}
	r.GetString();
	DWORD nSwFiles = r.GetDword();
	for ( DWORD i=0; i<nSwFiles && !r.m_bError; ++i )
	{
		r.GetString();
		SkipFileInfo ( r );
	}
	if ( r.GetByte() )
	{
		DWORD n = r.GetDword();
		for ( DWORD i=0; i<n && !r.m_bError; ++i )
			r.GetString();
	}
	DWORD nWfFiles = r.GetDword();
	for ( DWORD i=0; i<nWfFiles && !r.m_bError; ++i )
	{
		r.GetString();
		SkipFileInfo ( r );
	}
	r.GetDword();
	h.m_bWordDict = r.GetByte();
	r.GetByte();
	r.GetString();

	h.m_iDocinfo = r.GetOffset();
	h.m_iDocinfoIndex = r.GetOffset();
	h.m_iMinMaxIndex = r.GetOffset();

	DWORD nRegexps = r.GetDword();
	for ( DWORD i=0; i<nRegexps && !r.m_bError; ++i )
		r.GetString();

	if ( r.m_bError )
	{
		sError = "failed to parse header (unexpected eof)";
		return false;
	}
	return true;
}


BYTE DoclistHintPack ( int64_t iDocs, int64_t iLen )
{
	// sphDoclistHintPack, src/sphinx.cpp:10864-10878
	if ( iDocs<DOCLIST_HINT_THRESH )
		return 0;
	int64_t iDelta = std::min ( std::max ( iLen-4*iDocs, (int64_t)0 ), 4*iDocs-1 );
	BYTE uHint = (BYTE)( 64*iDelta/iDocs );
	while ( uHint<255 && ( iDocs*uHint/64 )<iDelta )
		uHint++;
	return uHint;
}


DictWriter_c::DictWriter_c ( int iSkiplistBlockSize, bool bCrc )
	: m_iSkiplistBlockSize ( iSkiplistBlockSize )
	, m_bCrc ( bCrc )
{
	m_tOut.PutByte ( 1 );	// CSphDictKeywords::DictBegin, src/sphinx.cpp:19382; CSphDiskDictTraits::DictBegin :18262-18267
}


void DictWriter_c::AddEntry ( const DictEntry_t & e )
{
	if ( m_bCrc )
	{
		// CSphDiskDictTraits::DictEntry, src/sphinx.cpp:18288-18330
		if ( ( m_iWords % SPH_WORDLIST_CHECKPOINT )==0 )
		{
			if ( m_iWords )
			{
				m_tOut.Zip ( 0 );	// indicate checkpoint
				m_tOut.Zip ( (uint64_t)( e.m_iDoclistOffset-m_iLastDoclistPos ) );	// store last length
			}
			m_uLastWordID = 0;
			m_iLastDoclistPos = 0;
			m_dCrcCheckpoints.push_back ( { e.m_uWordID, m_tOut.Pos() } );
		}
		m_tOut.Zip ( e.m_uWordID-m_uLastWordID );
		m_tOut.Zip ( (uint64_t)( e.m_iDoclistOffset-m_iLastDoclistPos ) );
		m_uLastWordID = e.m_uWordID;
		m_iLastDoclistPos = e.m_iDoclistOffset;
		m_tOut.Zip ( (uint64_t)e.m_iDocs );
		m_tOut.Zip ( (uint64_t)e.m_iHits );
		if ( e.m_iDocs>m_iSkiplistBlockSize )
			m_tOut.Zip ( (uint64_t)e.m_iSkiplistOffset );
		m_iWords++;
		return;
	}

	// src/sphinx.cpp:19468-19511
	if ( ( m_iWords % SPH_WORDLIST_CHECKPOINT )==0 )
	{
		if ( m_iWords )
		{
			m_tOut.Zip ( 0 );
			m_tOut.Zip ( 0 );
		}
		m_dCheckpoints.push_back ( { e.m_sKeyword, m_tOut.Pos() } );
		m_sLast.clear();
	}
	m_iWords++;

	// CSphKeywordDeltaWriter::PutDelta, src/sphinxint.h:1624-1657
	const std::string & w = e.m_sKeyword;
	int iLen = (int)w.size();
	int iMatch = 0;
	int iMinLen = std::min ( (int)m_sLast.size(), iLen );
	while ( iMatch<iMinLen && iMatch<255 && m_sLast[iMatch]==w[iMatch] )
		iMatch++;
	BYTE iDelta = (BYTE)( iLen-iMatch );
	m_sLast = w;
	if ( iDelta<=8 && iMatch<=15 )
		m_tOut.PutByte ( (BYTE)( 0x80 + ( ( iDelta-1 )<<4 ) + iMatch ) );
	else
	{
		m_tOut.PutByte ( iDelta );
		m_tOut.PutByte ( (BYTE)iMatch );
	}
	m_tOut.PutBytes ( w.data()+iMatch, iDelta );

	m_tOut.Zip ( (uint64_t)e.m_iDoclistOffset );
	m_tOut.Zip ( (uint64_t)e.m_iDocs );
	m_tOut.Zip ( (uint64_t)e.m_iHits );
	BYTE uHint = DoclistHintPack ( e.m_iDocs, e.m_iDoclistLength );
	if ( uHint )
		m_tOut.PutByte ( uHint );
	if ( e.m_iDocs>m_iSkiplistBlockSize )
		m_tOut.Zip ( (uint64_t)e.m_iSkiplistOffset );
}


void DictWriter_c::Finish ( IndexHeader_t & h, int64_t iDoclistEnd )
{
	if ( m_bCrc )
	{
		// DictEndEntries + DictEnd, src/sphinx.cpp:18269-18286, 18332-18337
		m_tOut.Zip ( 0 );
		m_tOut.Zip ( (uint64_t)( iDoclistEnd-m_iLastDoclistPos ) );
		h.m_iDictCheckpointsOffset = m_tOut.Pos();
		h.m_iDictCheckpoints = (DWORD)m_dCrcCheckpoints.size();
		for ( const auto & c : m_dCrcCheckpoints )
		{
			m_tOut.PutOffset ( (int64_t)c.first );
			m_tOut.PutOffset ( c.second );
		}
		h.m_iInfixCodepointBytes = 0;
		h.m_iInfixBlocksOffset = 0;
		h.m_iInfixBlocksWordsSize = 0;
		return;
	}
	// src/sphinx.cpp:19537-19576
	m_tOut.Zip ( 0 );
	m_tOut.Zip ( 0 );

	h.m_iDictCheckpointsOffset = m_tOut.Pos();
	h.m_iDictCheckpoints = (DWORD)m_dCheckpoints.size();
	for ( const auto & c : m_dCheckpoints )
	{
		m_tOut.PutDword ( (DWORD)c.m_sWord.size() );
		m_tOut.PutBytes ( c.m_sWord.data(), c.m_sWord.size() );
		m_tOut.PutOffset ( c.m_iOffset );
	}
	h.m_iInfixCodepointBytes = 0;
	h.m_iInfixBlocksOffset = 0;
	h.m_iInfixBlocksWordsSize = 0;

	m_tOut.PutBytes ( "dict-header", 11 );
	m_tOut.Zip ( h.m_iDictCheckpoints );
	m_tOut.Zip ( (uint64_t)h.m_iDictCheckpointsOffset );
	m_tOut.Zip ( h.m_iInfixCodepointBytes );
	m_tOut.Zip ( h.m_iInfixBlocksOffset );
}


bool ReadDictionary ( const BYTE * pSpi, size_t iLen, const IndexHeader_t & h, std::vector<DictEntry_t> & dOut, std::string & sError )
{
	dOut.clear();
	if ( !h.m_iDictCheckpoints )
		return true;
	if ( h.m_iDictCheckpointsOffset<=0 || (size_t)h.m_iDictCheckpointsOffset>iLen )
	{
		sError = "dictionary checkpoints offset out of bounds";
		return false;
	}
	if ( !h.m_bWordDict )
	{
		// dict=crc: checkpoints {u64 word id, u64 offset} (CWordlist::Preread -> CheckpointReader_c for crc dictionaries,
		// src/indexformat.cpp:354-410); chunks of delta-coded entries, closed by a zero id delta + the last doclist's length
		// (CWordlist::GetWord, :425-473)
		ByteReader_t rCp ( pSpi+h.m_iDictCheckpointsOffset, iLen-h.m_iDictCheckpointsOffset );
		const int iBlk = (int)h.m_iSkiplistBlockSize;
		for ( DWORD i=0; i<h.m_iDictCheckpoints; ++i )
		{
			rCp.GetOffset();	// the chunk's first word id (the entries repeat it)
			const int64_t iOff = rCp.GetOffset();
			if ( rCp.m_bError || iOff<=0 || iOff>=h.m_iDictCheckpointsOffset )
			{
				sError = "dictionary checkpoint out of bounds";
				return false;
			}
			ByteReader_t r ( pSpi+iOff, (size_t)( h.m_iDictCheckpointsOffset-iOff ) );
			uint64_t uLastID = 0;
			int64_t iLastOff = 0;
			while ( true )
			{
				const uint64_t uDelta = r.Unzip();
				if ( !uDelta || r.m_bError )
					break;
				DictEntry_t e;
				uLastID += uDelta;
				iLastOff += (int64_t)r.Unzip();
				e.m_uWordID = uLastID;
				e.m_sKeyword = CrcDictKey ( uLastID );
				e.m_iDoclistOffset = iLastOff;
				e.m_iDocs = (int)r.Unzip();
				e.m_iHits = (int)r.Unzip();
				if ( e.m_iDocs>iBlk )
					e.m_iSkiplistOffset = (int64_t)r.Unzip();
				if ( r.m_bError )
					break;
				dOut.push_back ( e );
			}
			if ( r.m_bError )
			{
				sError = "dictionary chunk truncated";
				return false;
			}
		}
		return true;
	}

	// checkpoint table: {u32 len, bytes, u64 offset}* (CWordlist::Preread, src/indexformat.cpp:331-344)
	ByteReader_t rCp ( pSpi+h.m_iDictCheckpointsOffset, iLen-h.m_iDictCheckpointsOffset );
	std::vector<int64_t> dCpOffsets;
	for ( DWORD i=0; i<h.m_iDictCheckpoints; ++i )
	{
		DWORD n = rCp.GetDword();
		if ( !rCp.Need ( n ) )
			break;
		rCp.m_p += n;
		dCpOffsets.push_back ( rCp.GetOffset() );
	}
	if ( rCp.m_bError )
	{
		sError = "dictionary checkpoints truncated";
		return false;
	}

	const int iBlk = (int)h.m_iSkiplistBlockSize;
	for ( int64_t iOff : dCpOffsets )
	{
		if ( iOff<=0 || iOff>=h.m_iDictCheckpointsOffset )
		{
			sError = "dictionary checkpoint offset out of bounds";
			return false;
		}
		// KeywordsBlockReader_c::UnpackWord, src/indexformat.cpp:641-691
		ByteReader_t r ( pSpi+iOff, (size_t)( h.m_iDictCheckpointsOffset-iOff ) );
		std::string sWord;
		while ( true )
		{
			BYTE uPack = r.GetByte();
			if ( !uPack || r.m_bError )
				break;
			int iMatch, iDelta;
			if ( uPack & 0x80 )
			{
				iDelta = ( ( uPack>>4 ) & 7 ) + 1;
				iMatch = uPack & 15;
			} else
			{
				iDelta = uPack & 127;
				iMatch = r.GetByte();
			}
			if ( iMatch>(int)sWord.size() || !r.Need ( iDelta ) )
			{
				sError = "corrupt keyword delta in dictionary";
				return false;
			}
			sWord.resize ( iMatch );
			sWord.append ( (const char*)r.m_p, iDelta );
			r.m_p += iDelta;

			DictEntry_t e;
			e.m_sKeyword = sWord;
			e.m_iDoclistOffset = (int64_t)r.Unzip();
			e.m_iDocs = (int)r.Unzip();
			e.m_iHits = (int)r.Unzip();
			if ( e.m_iDocs>=DOCLIST_HINT_THRESH )
				r.GetByte();
			if ( e.m_iDocs>iBlk )
				e.m_iSkiplistOffset = (int64_t)r.Unzip();
			dOut.push_back ( e );
		}
		if ( r.m_bError )
		{
			sError = "dictionary block truncated";
			return false;
		}
	}
	return true;
}

} // namespace mgpu
