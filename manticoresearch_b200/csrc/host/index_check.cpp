// Index consistency check of the files the query path loads (.sph .spi .spd .spp .spe .spa .spm), host only.
//
// Restates the checks of DiskIndexChecker_c (src/indexcheck.cpp: CheckDictionary :461-681, CheckDocs :684-983, DebugCheck_Attributes
// :149-256, DebugCheck_DeadRowMap :259-267, CheckDocids :1292-1317) for plain indexes of format v57..v62 with inline or plain hit
// format; what searchd's `indextool --check` would say about an index before mgpu_index_open uploads it.  Not checked: blobs, docstore,
// kill lists, the docid lookup (.spt) and the min-max block index, none of which the query path reads.
#include "index_format.h"
#include "../../../include/mgpu.h"

#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <unordered_set>

namespace mgpu
{
namespace
{

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
			if ( p==MAP_FAILED ) { close ( fd ); m_iLen = 0; return false; }
			m_p = (const BYTE *)p;
		}
		close ( fd );
		return true;
	}
	~Mapped_t() { if ( m_p ) munmap ( (void*)m_p, m_iLen ); }
};

/// DebugCheckError_c (src/indexcheck.cpp:31-99): counts failures, keeps the first messages
struct Reporter_t
{
	int64_t		m_nFails = 0;
	std::string	m_sLog;
	void Fail ( const char * sFmt, ... ) __attribute__ ( ( format ( printf, 2, 3 ) ) )
	{
		++m_nFails;
		if ( m_nFails>64 )
			return;
		char sBuf[512];
		va_list ap;
		va_start ( ap, sFmt );
		vsnprintf ( sBuf, sizeof(sBuf), sFmt, ap );
		va_end ( ap );
		m_sLog += "FAILED, ";
		m_sLog += sBuf;
		m_sLog += "\n";
	}
};

static std::string WordName ( const DictEntry_t & e, bool bWordDict )
{
	if ( bWordDict )
		return e.m_sKeyword;
	char s[32];
	snprintf ( s, sizeof(s), "wordid=%llu", (unsigned long long)e.m_uWordID );
	return s;
}

struct SkipEntry_t { uint64_t m_uBaseRowIDPlus1, m_uOffset, m_uHitPos; };

/// CheckDocs, src/indexcheck.cpp:684-983: every doclist decoded end to end, its hitlists walked, the skiplist recomputed
static void CheckDocs ( const IndexHeader_t & h, const std::vector<DictEntry_t> & dDict, const Mapped_t & tSpd, const Mapped_t & tSpp, const Mapped_t & tSpe, Reporter_t & R )
{
	const bool bInline = ( h.m_eHitFormat==SPH_HIT_FORMAT_INLINE );
	const int iBlk = (int)h.m_iSkiplistBlockSize;
	const int nFields = (int)h.m_dFields.size();
	const BYTE * pSpdEnd = tSpd.m_p + tSpd.m_iLen;
	const BYTE * pSppEnd = tSpp.m_p + tSpp.m_iLen;
	int64_t iExpectedDoclist = 1;		// after the dummy byte; doclists follow each other in dictionary order
	uint64_t uExpectedHitlist = 1;
	int64_t iLastSkiplist = 0;
	std::vector<SkipEntry_t> dSkips;

	for ( const DictEntry_t & e : dDict )
	{
		const std::string sWord = WordName ( e, h.m_bWordDict!=0 );
		if ( e.m_iDoclistOffset!=iExpectedDoclist && !h.m_bWordDict )	// (keywords dictionaries may order their doclists differently, :787-791)
			R.Fail ( "unexpected doclist offset (%s, dict=%lld, expected=%lld)", sWord.c_str(), (long long)e.m_iDoclistOffset, (long long)iExpectedDoclist );
		if ( e.m_iDoclistOffset<=0 || e.m_iDoclistOffset>=(int64_t)tSpd.m_iLen )
		{
			R.Fail ( "unexpected doclist offset, off the file (%s, doclistpos=%lld, doclistsize=%lld)", sWord.c_str(), (long long)e.m_iDoclistOffset, (long long)tSpd.m_iLen );
			return;		// cannot resynchronise
		}

		const BYTE * p = tSpd.m_p + e.m_iDoclistOffset;
		uint64_t uRow = (uint64_t)-1;		// INVALID_ROWID: the first delta is rowid+1
		uint64_t uHitPos = 0;				// inline: m_uHitPosition; plain: m_iHitlistPos (both running sums)
		int64_t nDocs = 0, nDoclistHits = 0, nHitlistHits = 0;
		dSkips.clear();
		bool bBroken = false;
		while ( true )
		{
			if ( p>=pSpdEnd )
			{
				R.Fail ( "doclist runs off the file (%s)", sWord.c_str() );
				bBroken = true;
				break;
			}
			const BYTE * pRecord = p;
			ByteReader_t r ( p, (size_t)( pSpdEnd-p ) );
			const uint64_t uDelta = r.Unzip();
			if ( !uDelta )
			{
				p = r.m_p;
				break;
			}
			if ( ( nDocs % iBlk )==0 )
				dSkips.push_back ( { uRow+1, (uint64_t)( pRecord-tSpd.m_p ), uHitPos } );
			uRow += uDelta;
			++nDocs;
			if ( uRow>=(uint64_t)h.m_iDocinfo )
				R.Fail ( "rowid out of bounds (%s, rowid=%llu)", sWord.c_str(), (unsigned long long)uRow );

			uint32_t uFields, uMatchHits;
			uint64_t uHitlistAt = 0;
			bool bInlinedOne = false;
			Hitpos_t uInlined = 0;
			if ( bInline )
			{
				uMatchHits = (uint32_t)r.Unzip();
				const uint32_t uFirst = (uint32_t)r.Unzip();
				if ( uMatchHits==1 )
				{
					const uint32_t uField = (uint32_t)r.Unzip();
					bInlinedOne = true;
					uInlined = uFirst | ( uField<<23 );
					const uint32_t iField = ( uField>>1 ) & 255;
					uFields = iField<32 ? ( 1u<<iField ) : 0;
				} else
				{
					uFields = uFirst;
					uHitPos += r.Unzip();
					uHitlistAt = uHitPos;
				}
			} else
			{
				uHitPos += r.Unzip();
				uHitlistAt = uHitPos;
				uFields = (uint32_t)r.Unzip();
				uMatchHits = (uint32_t)r.Unzip();
			}
			if ( r.m_bError )
			{
				R.Fail ( "doclist record truncated (%s, rowid=%llu)", sWord.c_str(), (unsigned long long)uRow );
				bBroken = true;
				break;
			}
			p = r.m_p;
			nDoclistHits += uMatchHits;
			if ( !uMatchHits )
				R.Fail ( "doc without hits (%s, rowid=%llu)", sWord.c_str(), (unsigned long long)uRow );

			// hits
			uint32_t uHitFields = 0;
			int64_t nDocHits = 0;
			Hitpos_t uLastHit = 0;
			auto fnHit = [&] ( Hitpos_t uHit )
			{
				if ( nDocHits && uLastHit>=uHit )
					R.Fail ( "hit entries sorting order decreased (%s, rowid=%llu, hit=%u, last=%u)", sWord.c_str(), (unsigned long long)uRow, uHit, uLastHit );
				if ( nDocHits && HITMAN::GetField ( uLastHit )==HITMAN::GetField ( uHit ) )
				{
					if ( HITMAN::GetPos ( uLastHit )>=HITMAN::GetPos ( uHit ) )
						R.Fail ( "hit decreased (%s, rowid=%llu, hit=%d, last=%d)", sWord.c_str(), (unsigned long long)uRow, HITMAN::GetPos ( uHit ), HITMAN::GetPos ( uLastHit ) );
					// (a second hit behind a field-end hit is only a warning there, :878-879)
				} else if ( nDocHits && HITMAN::GetField ( uLastHit )>HITMAN::GetField ( uHit ) )
					R.Fail ( "hit field decreased (%s, rowid=%llu, hit field=%d, last field=%d)", sWord.c_str(), (unsigned long long)uRow, HITMAN::GetField ( uHit ), HITMAN::GetField ( uLastHit ) );
				const int iField = HITMAN::GetField ( uHit );
				if ( iField>=nFields )
					R.Fail ( "hit field out of schema (%s, rowid=%llu, field=%d)", sWord.c_str(), (unsigned long long)uRow, iField );
				else if ( iField<32 )
					uHitFields |= 1u<<iField;
				uLastHit = uHit;
				++nDocHits;
			};
			if ( bInlinedOne )
				fnHit ( uInlined );
			else
			{
				if ( uHitlistAt!=uExpectedHitlist && !h.m_bWordDict )
					R.Fail ( "unexpected hitlist offset (%s, rowid=%llu, expected=%llu, actual=%llu)", sWord.c_str(), (unsigned long long)uRow,
						(unsigned long long)uExpectedHitlist, (unsigned long long)uHitlistAt );
				if ( uHitlistAt==0 || uHitlistAt>=tSpp.m_iLen )
				{
					R.Fail ( "hitlist offset off the file (%s, rowid=%llu, offset=%llu)", sWord.c_str(), (unsigned long long)uRow, (unsigned long long)uHitlistAt );
					bBroken = true;
					break;
				}
				ByteReader_t rh ( tSpp.m_p+uHitlistAt, (size_t)( pSppEnd-( tSpp.m_p+uHitlistAt ) ) );
				Hitpos_t uHit = 0;
				while ( true )
				{
					const uint32_t d = (uint32_t)rh.Unzip();
					if ( !d || rh.m_bError )
						break;
					uHit += d;
					fnHit ( uHit );
				}
				if ( rh.m_bError )
					R.Fail ( "hitlist runs off the file (%s, rowid=%llu)", sWord.c_str(), (unsigned long long)uRow );
				uExpectedHitlist = (uint64_t)( rh.m_p-tSpp.m_p );
			}
			nHitlistHits += nDocHits;
			if ( nDocHits!=(int64_t)uMatchHits )
				R.Fail ( "doc hit count mismatch (%s, rowid=%llu, doclist=%u, hitlist=%lld)", sWord.c_str(), (unsigned long long)uRow, uMatchHits, (long long)nDocHits );
			if ( nFields<=32 && uFields!=uHitFields )
				R.Fail ( "field mask mismatch (%s, rowid=%llu, doclist=0x%x, hitlist=0x%x)", sWord.c_str(), (unsigned long long)uRow, uFields, uHitFields );
		}
		if ( bBroken )
			return;
		iExpectedDoclist = (int64_t)( p-tSpd.m_p );

		if ( nDocs!=e.m_iDocs )
			R.Fail ( "doc count mismatch (%s, dict=%d, doclist=%lld)", sWord.c_str(), e.m_iDocs, (long long)nDocs );
		if ( nDoclistHits!=e.m_iHits || nHitlistHits!=e.m_iHits )
			R.Fail ( "hit count mismatch (%s, dict=%d, doclist=%lld, hitlist=%lld)", sWord.c_str(), e.m_iHits, (long long)nDoclistHits, (long long)nHitlistHits );

		// skiplist: stored for docs > block size, entry 0 implicit, deltas against { +block size, +4*block size, +0 } (src/sphinx.cpp:13056-13073)
		if ( e.m_iDocs>iBlk )
		{
			if ( e.m_iSkiplistOffset<=0 || e.m_iSkiplistOffset>=(int64_t)tSpe.m_iLen )
			{
				R.Fail ( "invalid skiplist offset (%s, off=%lld, max=%lld)", sWord.c_str(), (long long)e.m_iSkiplistOffset, (long long)tSpe.m_iLen );
				continue;
			}
			if ( e.m_iSkiplistOffset<=iLastSkiplist )
				R.Fail ( "descending skiplist pos (last=%lld, cur=%lld, %s)", (long long)iLastSkiplist, (long long)e.m_iSkiplistOffset, sWord.c_str() );
			iLastSkiplist = e.m_iSkiplistOffset;
			ByteReader_t rs ( tSpe.m_p+e.m_iSkiplistOffset, tSpe.m_iLen-(size_t)e.m_iSkiplistOffset );
			SkipEntry_t t = { 0, (uint64_t)e.m_iDoclistOffset, 0 };
			for ( size_t i=1; i<dSkips.size(); ++i )
			{
				t.m_uBaseRowIDPlus1 += (uint64_t)iBlk + rs.Unzip();
				t.m_uOffset += 4*(uint64_t)iBlk + rs.Unzip();
				t.m_uHitPos += rs.Unzip();
				if ( rs.m_bError )
				{
					R.Fail ( "skiplist reading error (%s, exp=%d, got=%d)", sWord.c_str(), (int)dSkips.size(), (int)i );
					break;
				}
				const SkipEntry_t & x = dSkips[i];
				if ( t.m_uBaseRowIDPlus1!=x.m_uBaseRowIDPlus1 || t.m_uOffset!=x.m_uOffset || t.m_uHitPos!=x.m_uHitPos )
				{
					R.Fail ( "skiplist entry %d mismatch (%s, exp={%llu, %llu, %llu}, got={%llu, %llu, %llu})", (int)i, sWord.c_str(),
						(unsigned long long)x.m_uBaseRowIDPlus1, (unsigned long long)x.m_uOffset, (unsigned long long)x.m_uHitPos,
						(unsigned long long)t.m_uBaseRowIDPlus1, (unsigned long long)t.m_uOffset, (unsigned long long)t.m_uHitPos );
					break;
				}
			}
		}
	}
}

/// CheckDictionary, src/indexcheck.cpp:461-681 (on the decoded entries: order, counts, offsets, checkpoint count)
static void CheckDictionary ( const IndexHeader_t & h, const std::vector<DictEntry_t> & dDict, Reporter_t & R )
{
	const int64_t nExpectedCp = ( (int64_t)dDict.size()+SPH_WORDLIST_CHECKPOINT-1 )/SPH_WORDLIST_CHECKPOINT;
	if ( (int64_t)h.m_iDictCheckpoints!=nExpectedCp )
		R.Fail ( "checkpoint count mismatch (read=%u, calc=%lld)", h.m_iDictCheckpoints, (long long)nExpectedCp );
	for ( size_t i=0; i<dDict.size(); ++i )
	{
		const DictEntry_t & e = dDict[i];
		const std::string sWord = WordName ( e, h.m_bWordDict!=0 );
		if ( h.m_bWordDict && e.m_sKeyword.empty() )
			R.Fail ( "empty word in dictionary (entry %d)", (int)i );
		if ( e.m_iDocs<=0 || e.m_iHits<=0 || e.m_iHits<e.m_iDocs )
			R.Fail ( "invalid docs/hits (%s, docs=%d, hits=%d)", sWord.c_str(), e.m_iDocs, e.m_iHits );
		if ( !i )
			continue;
		const DictEntry_t & tPrev = dDict[i-1];
		if ( h.m_bWordDict ? strcmp ( tPrev.m_sKeyword.c_str(), e.m_sKeyword.c_str() )>=0 : tPrev.m_uWordID>=e.m_uWordID )
			R.Fail ( h.m_bWordDict ? "word order decreased (%s, prev=%s)" : "wordid decreased (%s, prev %s)", sWord.c_str(), WordName ( tPrev, h.m_bWordDict!=0 ).c_str() );
		if ( e.m_iDoclistOffset<=tPrev.m_iDoclistOffset )
			R.Fail ( "doclist offset decreased (%s)", sWord.c_str() );
	}
}

} // namespace
} // namespace mgpu

static int IndexCheckImpl ( const char * path_prefix, int64_t * n_failures, char * report, int report_len );

extern "C" int mgpu_index_check ( const char * path_prefix, int64_t * n_failures, char * report, int report_len )
{
	try
	{
		return IndexCheckImpl ( path_prefix, n_failures, report, report_len );
	} catch ( ... )
	{
		return MGPU_E_NOMEM;	// no exception crosses the ABI
	}
}

static int IndexCheckImpl ( const char * path_prefix, int64_t * n_failures, char * report, int report_len )
{
	using namespace mgpu;
	if ( n_failures )
		*n_failures = 0;
	if ( report && report_len>0 )
		report[0] = '\0';
	if ( !path_prefix )
		return MGPU_E_BAD_QUERY;
	Reporter_t R;
	auto fnFinish = [&] ( int iRes )
	{
		if ( n_failures )
			*n_failures = R.m_nFails;
		if ( report && report_len>0 )
		{
			strncpy ( report, R.m_sLog.c_str(), (size_t)report_len-1 );
			report[report_len-1] = '\0';
		}
		return iRes;
	};

	const std::string sPrefix ( path_prefix );
	Mapped_t tSph, tSpi, tSpd, tSpp, tSpe, tSpa, tSpm;
	struct { Mapped_t * m_p; const char * m_sExt; const char * m_sWhat; } dFiles[] = {
		{ &tSph, ".sph", "header" }, { &tSpi, ".spi", "dictionary" }, { &tSpd, ".spd", "doclist" }, { &tSpp, ".spp", "hitlist" },
		{ &tSpe, ".spe", "skiplist" }, { &tSpa, ".spa", "attributes" }, { &tSpm, ".spm", "dead-row map" } };
	for ( auto & f : dFiles )
		if ( !f.m_p->Map ( sPrefix+f.m_sExt ) )
		{
			R.Fail ( "unable to open %s: %s%s", f.m_sWhat, path_prefix, f.m_sExt );
			return fnFinish ( MGPU_E_IO );
		}

	IndexHeader_t h;
	std::string sError;
	if ( !ReadHeader ( tSph.m_p, tSph.m_iLen, h, sError ) )
	{
		R.Fail ( "error reading index header: %s", sError.c_str() );
		return fnFinish ( MGPU_E_FORMAT );
	}
	if ( h.m_eHitless!=SPH_HITLESS_NONE )
	{
		R.Fail ( "hitless indexes are not supported" );
		return fnFinish ( MGPU_E_FORMAT );
	}
	if ( h.m_iDocinfo<0 || h.m_iDocinfo>=(int64_t)0xFFFFFFFFll || h.m_dAttrs.size()>4096 )
	{
		R.Fail ( "implausible row count %lld or attribute count %d in the header", (long long)h.m_iDocinfo, (int)h.m_dAttrs.size() );
		return fnFinish ( MGPU_E_FORMAT );
	}
	if ( (int)h.m_iSkiplistBlockSize<=0 )
	{
		R.Fail ( "invalid skiplist block size %u", h.m_iSkiplistBlockSize );
		return fnFinish ( MGPU_E_FORMAT );
	}

	// schema, DebugCheck_Attributes :149-170 + DebugCheckSchema_T :1363-1380
	if ( h.m_dAttrs.empty() )
		R.Fail ( "no attributes in schema; schema should at least have 'id' attr" );
	else
	{
		if ( h.m_dAttrs[0].m_sName!="id" )
			R.Fail ( "first attribute in schema should be 'id'" );
		if ( h.m_dAttrs[0].m_iBitCount!=64 )
			R.Fail ( "id attribute should be BIGINT" );
		std::unordered_set<std::string> hNames;
		for ( const auto & a : h.m_dAttrs )
			if ( !hNames.insert ( a.m_sName ).second )
				R.Fail ( "duplicate attributes name %s", a.m_sName.c_str() );
	}

	// attribute rows + dead-row map sizes (:198-200, :259-267)
	const int iStride = h.RowStride();
	const int64_t iRowItems = (int64_t)( tSpa.m_iLen/4 );
	const int64_t iExpectedItems = h.m_iDocinfo*iStride + ( h.m_iMinMaxIndex>0 ? iRowItems-h.m_iMinMaxIndex : 0 );
	if ( tSpa.m_iLen % 4 || iRowItems<h.m_iDocinfo*iStride || ( h.m_iMinMaxIndex>0 && ( h.m_iMinMaxIndex!=h.m_iDocinfo*iStride || iRowItems!=iExpectedItems ) ) )
		R.Fail ( "rowitems count mismatch (expected=%lld, loaded=%lld)", (long long)( h.m_iDocinfo*iStride ), (long long)iRowItems );
	const int64_t iExpectedSpm = ( ( h.m_iDocinfo+31 )/32 )*4;
	if ( (int64_t)tSpm.m_iLen!=iExpectedSpm )
		R.Fail ( "unexpected dead row map: %lld, expected: %lld bytes", (long long)tSpm.m_iLen, (long long)iExpectedSpm );

	// duplicate document ids, CheckDocids :1292-1317
	if ( !h.m_dAttrs.empty() && h.m_dAttrs[0].m_iBitCount==64 && iRowItems>=h.m_iDocinfo*iStride && iStride>=2 )
	{
		std::vector<std::pair<int64_t,uint32_t>> dIds ( (size_t)h.m_iDocinfo );
		const DWORD * pRows = (const DWORD *)tSpa.m_p;
		for ( int64_t i=0; i<h.m_iDocinfo; ++i )
		{
			int64_t iId;
			memcpy ( &iId, pRows+i*iStride, 8 );
			dIds[(size_t)i] = { iId, (uint32_t)i };
		}
		std::sort ( dIds.begin(), dIds.end() );
		for ( size_t i=1; i<dIds.size(); ++i )
			if ( dIds[i].first==dIds[i-1].first )
				R.Fail ( "duplicate of docid %lld found at rows %u %u", (long long)dIds[i].first, dIds[i-1].second, dIds[i].second );
	}

	std::vector<DictEntry_t> dDict;
	if ( !ReadDictionary ( tSpi.m_p, tSpi.m_iLen, h, dDict, sError ) )
	{
		R.Fail ( "dictionary: %s", sError.c_str() );
		return fnFinish ( MGPU_OK );
	}
	for ( auto * p : { &tSpd, &tSpp, &tSpe } )
		if ( !p->m_iLen || p->m_p[0]!=1 )
			R.Fail ( "data file does not start with the dummy byte" );
	CheckDictionary ( h, dDict, R );
	CheckDocs ( h, dDict, tSpd, tSpp, tSpe, R );
	return fnFinish ( MGPU_OK );
}
