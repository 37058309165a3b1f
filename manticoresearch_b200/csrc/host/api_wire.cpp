// Wire responder for the binary SphinxAPI `search` and `keywords` commands, on top of the C ABI.
//
// Restates the server side of SEARCHD_COMMAND_SEARCH for one local index: the packet header (APIHeader / APIAnswer, src/searchdaemon.cpp),
// ParseSearchQuery + ParseSearchFilter (src/searchd.cpp:2201-2560), the legacy sort modes and the extended sort clause
// (sphCreateQueue / sphParseSortClause, src/sortsetup.cpp), per-field weights (CSphQueryContext::BindWeights, src/sphinx.cpp:13903-13947)
// and SendResult / SendSchema (src/searchd.cpp:3340-3510): request bytes in, reply bytes out, the search itself through
// mgpu_parse_query + mgpu_search_batch.  What the hot path has no counterpart for is answered per query with SEARCHD_ERROR, as searchd
// answers a query it cannot run: group-by, expression rankers and sorts, geo anchors, select lists other than "*", string / float filters,
// outer order, query token filters.  No sockets here: the embedding daemon owns the connection.
#include "index_format.h"
#include "../../../include/mgpu.h"

#include <fcntl.h>
#include <stdio.h>
#include <strings.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <map>
#include <memory>

namespace mgpu
{
void TokenizePlain ( const mgpu_parser_settings & s, const char * sText, std::vector<std::pair<std::string,int>> & dOut );	// query_parser.cpp

namespace
{

enum { SEARCHD_COMMAND_SEARCH = 0, SEARCHD_COMMAND_KEYWORDS = 3 };
enum { VER_COMMAND_KEYWORDS = 0x101 };	// src/searchdaemon.h:189
enum { SEARCHD_OK = 0, SEARCHD_ERROR = 1, SEARCHD_RETRY = 2, SEARCHD_WARNING = 3 };
enum { VER_COMMAND_SEARCH = 0x121 };	// src/searchdaemon.h:186
enum { SPH_SORT_RELEVANCE = 0, SPH_SORT_ATTR_DESC = 1, SPH_SORT_ATTR_ASC = 2, SPH_SORT_TIME_SEGMENTS = 3, SPH_SORT_EXTENDED = 4, SPH_SORT_EXPR = 5 };
enum { SPH_FILTER_VALUES = 0, SPH_FILTER_RANGE = 1, SPH_FILTER_FLOATRANGE = 2, SPH_FILTER_STRING = 3, SPH_FILTER_NULL = 4, SPH_FILTER_USERVAR = 5, SPH_FILTER_STRING_LIST = 6 };
enum { QFLAG_PLAIN_IDF = 1u<<4, QFLAG_NORMALIZED_TF = 1u<<6, QFLAG_MAX_PREDICTED_TIME = 1u<<2 };	// src/searchdaemon.h (QueryFlags_e)
enum { SPH_RANK_EXPR = 8, SPH_RANK_EXPORT = 9 };
enum { SPH_ATTR_BOOL = 4, SPH_ATTR_FLOAT = 5 };

/// InputBuffer_c (src/searchdaemon.h): network byte order
struct NetReader_t
{
	const BYTE * m_p;
	const BYTE * m_pEnd;
	bool m_bError = false;
	NetReader_t ( const void * p, size_t n ) : m_p ( (const BYTE*)p ), m_pEnd ( (const BYTE*)p+n ) {}
	bool Need ( size_t n )			{ if ( (size_t)( m_pEnd-m_p )<n ) { m_bError = true; m_p = m_pEnd; return false; } return true; }
	uint16_t GetWord ()				{ if ( !Need(2) ) return 0; uint16_t v = (uint16_t)( ( m_p[0]<<8 ) | m_p[1] ); m_p += 2; return v; }
	uint32_t GetDword ()			{ if ( !Need(4) ) return 0; uint32_t v = ( (uint32_t)m_p[0]<<24 ) | ( (uint32_t)m_p[1]<<16 ) | ( (uint32_t)m_p[2]<<8 ) | m_p[3]; m_p += 4; return v; }
	int GetInt ()					{ return (int)GetDword(); }
	uint64_t GetUint64 ()			{ uint64_t hi = GetDword(); uint64_t lo = GetDword(); return ( hi<<32 ) | lo; }
	float GetFloat ()				{ uint32_t u = GetDword(); float f; memcpy ( &f, &u, 4 ); return f; }
	std::string GetString ()
	{
		uint32_t n = GetDword();
		if ( m_bError || !Need ( n ) )
			return std::string();
		std::string s ( (const char*)m_p, n );
		m_p += n;
		return s;
	}
};

/// ISphOutputBuffer
struct NetWriter_t
{
	std::vector<BYTE> m_d;
	void SendWord ( uint16_t v )		{ m_d.push_back ( (BYTE)( v>>8 ) ); m_d.push_back ( (BYTE)v ); }
	void SendDword ( uint32_t v )		{ for ( int s=24; s>=0; s-=8 ) m_d.push_back ( (BYTE)( v>>s ) ); }
	void SendInt ( int v )				{ SendDword ( (uint32_t)v ); }
	void SendUint64 ( uint64_t v )		{ SendDword ( (uint32_t)( v>>32 ) ); SendDword ( (uint32_t)v ); }
	void SendAsDword ( int64_t v )		{ SendDword ( (uint32_t)std::min<int64_t> ( std::max<int64_t> ( v, 0 ), 0xFFFFFFFFll ) ); }	// ISphOutputBuffer::SendAsDword clamps
	void SendString ( const std::string & s )	{ SendDword ( (uint32_t)s.size() ); m_d.insert ( m_d.end(), s.begin(), s.end() ); }
	void PatchDword ( size_t iAt, uint32_t v )	{ for ( int i=0; i<4; ++i ) m_d[iAt+i] = (BYTE)( v>>( 24-8*i ) ); }
};

struct ApiFilter_t
{
	std::string				m_sAttr;
	int						m_eType = SPH_FILTER_VALUES;
	std::vector<int64_t>	m_dValues;
	int64_t					m_iMin = 0, m_iMax = 0;
	bool					m_bExclude = false;
	std::string				m_sUnsupported;
};

/// the CSphQuery members ParseSearchQuery fills
struct ApiQuery_t
{
	uint32_t	m_uFlags = 0;
	int			m_iOffset = 0, m_iLimit = 20, m_eMode = 0, m_eRanker = 0, m_eSort = 0;
	std::string	m_sRankerExpr, m_sSortBy, m_sRawQuery, m_sIndexes, m_sGroupBy, m_sGroupSortBy, m_sGroupDistinct, m_sComment, m_sSelect, m_sOuterOrderBy;
	std::vector<uint32_t> m_dWeights;
	std::vector<ApiFilter_t> m_dFilters;
	int			m_eGroupFunc = 0, m_iMaxMatches = 1000, m_iCutoff = 0;
	bool		m_bGeoAnchor = false, m_bHasOuter = false, m_bTokenFilter = false;
	std::vector<std::pair<std::string,int>> m_dIndexWeights, m_dFieldWeights;
	uint32_t	m_uMaxQueryMsec = 0;
	std::string	m_sError;		// set while parsing: the query is answered with SEARCHD_ERROR
};

static std::string ToLower ( std::string s )
{
	for ( char & c : s )
		c = (char)tolower ( (unsigned char)c );
	return s;
}

/// ParseSearchFilter, src/searchd.cpp:2201-2318 (client protocol: no master extensions)
static bool ParseFilter ( NetReader_t & r, ApiFilter_t & f )
{
	f.m_sAttr = ToLower ( r.GetString() );
	f.m_eType = r.GetInt();
	switch ( f.m_eType )
	{
	case SPH_FILTER_RANGE:
		f.m_iMin = (int64_t)r.GetUint64();
		f.m_iMax = (int64_t)r.GetUint64();
		break;
	case SPH_FILTER_FLOATRANGE:
		r.GetFloat(); r.GetFloat();
		f.m_sUnsupported = "float range filters are not supported";
		break;
	case SPH_FILTER_VALUES:
		{
			int n = r.GetInt();
			if ( n<0 || (size_t)n*8>(size_t)( r.m_pEnd-r.m_p ) )
			{
				r.m_bError = true;
				return false;
			}
			f.m_dValues.resize ( n );
			for ( auto & v : f.m_dValues )
				v = (int64_t)r.GetUint64();
			std::sort ( f.m_dValues.begin(), f.m_dValues.end() );	// FixupQuerySettings
		}
		break;
	case SPH_FILTER_STRING:
		r.GetString();
		f.m_sUnsupported = "string filters are not supported";
		break;
	case SPH_FILTER_STRING_LIST:
		{
			int n = r.GetInt();
			for ( int i=0; i<n && !r.m_bError; ++i )
				r.GetString();
			f.m_sUnsupported = "string filters are not supported";
		}
		break;
	case SPH_FILTER_NULL:
		r.m_p += r.Need(1) ? 1 : 0;
		f.m_sUnsupported = "null filters are not supported";
		break;
	default:
		return false;	// "unknown filter type"
	}
	f.m_bExclude = r.GetDword()!=0;
	return !r.m_bError;
}

/// ParseSearchQuery, src/searchd.cpp:2320-2560, for a client (uMasterVer==0)
static bool ParseQuery ( NetReader_t & r, ApiQuery_t & q, uint16_t uVer, std::string & sFatal )
{
	if ( uVer>=0x11B )
		q.m_uFlags = r.GetDword();
	q.m_iOffset = r.GetInt();
	q.m_iLimit = r.GetInt();
	q.m_eMode = r.GetInt();
	q.m_eRanker = r.GetInt();
	if ( q.m_eRanker==SPH_RANK_EXPR || q.m_eRanker==SPH_RANK_EXPORT )
		q.m_sRankerExpr = r.GetString();
	q.m_eSort = r.GetInt();
	q.m_sSortBy = ToLower ( r.GetString() );
	q.m_sRawQuery = r.GetString();
	{
		int n = r.GetInt();
		if ( n<0 || n>256 )
		{
			sFatal = "invalid weight count " + std::to_string ( n ) + " (should be in 0..256 range)";
			return false;
		}
		q.m_dWeights.resize ( n );
		for ( auto & w : q.m_dWeights )
			w = r.GetDword();
	}
	q.m_sIndexes = r.GetString();
	const bool bId64 = r.GetInt()!=0;
	int64_t iMinId = bId64 ? (int64_t)r.GetUint64() : (int64_t)r.GetDword();
	int64_t iMaxId = bId64 ? (int64_t)r.GetUint64() : (int64_t)r.GetDword();
	if ( iMaxId==0 || (uint64_t)iMaxId==UINT64_MAX )
		iMaxId = INT64_MAX;

	int nFilters = r.GetInt();
	if ( nFilters<0 || nFilters>256 )
	{
		sFatal = "too many attribute filters (req=" + std::to_string ( nFilters ) + ", max=256)";
		return false;
	}
	q.m_dFilters.resize ( nFilters );
	for ( auto & f : q.m_dFilters )
		if ( !ParseFilter ( r, f ) )
		{
			sFatal = r.m_bError ? "invalid or truncated request" : "unknown filter type (type-id=" + std::to_string ( f.m_eType ) + ")";
			return false;
		}
	if ( iMinId!=0 || iMaxId!=INT64_MAX )
	{
		ApiFilter_t f;
		f.m_sAttr = "id";
		f.m_eType = SPH_FILTER_RANGE;
		f.m_iMin = iMinId;
		f.m_iMax = iMaxId;
		q.m_dFilters.push_back ( f );
	}

	q.m_eGroupFunc = (int)r.GetDword();
	q.m_sGroupBy = ToLower ( r.GetString() );
	q.m_iMaxMatches = r.GetInt();
	q.m_sGroupSortBy = r.GetString();
	q.m_iCutoff = r.GetInt();
	r.GetInt(); r.GetInt();		// retry count / delay: agents only
	q.m_sGroupDistinct = r.GetString();
	q.m_bGeoAnchor = r.GetInt()!=0;
	if ( q.m_bGeoAnchor )
	{
		r.GetString(); r.GetString(); r.GetFloat(); r.GetFloat();
	}
	{
		int n = r.GetInt();
		for ( int i=0; i<n && !r.m_bError; ++i )
		{
			std::string s = r.GetString();
			q.m_dIndexWeights.push_back ( { s, r.GetInt() } );
		}
	}
	q.m_uMaxQueryMsec = r.GetDword();
	{
		int n = r.GetInt();
		for ( int i=0; i<n && !r.m_bError; ++i )
		{
			std::string s = r.GetString();
			q.m_dFieldWeights.push_back ( { s, r.GetInt() } );
		}
	}
	q.m_sComment = r.GetString();
	if ( r.GetInt()>0 )
	{
		sFatal = "overrides are now deprecated";
		return false;
	}
	q.m_sSelect = r.GetString();
	if ( q.m_sSelect.empty() )
		q.m_sSelect = "*";
	if ( uVer>=0x11B && ( q.m_uFlags & QFLAG_MAX_PREDICTED_TIME ) )
		r.GetInt();
	if ( uVer>=0x11D )
	{
		q.m_sOuterOrderBy = r.GetString();
		r.GetDword(); r.GetDword();
		q.m_bHasOuter = r.GetInt()!=0;
	}
	if ( uVer>=0x120 )
	{
		std::string sLib = r.GetString(), sName = r.GetString();
		r.GetString();
		q.m_bTokenFilter = !sLib.empty() || !sName.empty();
	}
	if ( uVer>=0x121 )
	{
		int n = r.GetInt();
		if ( n<0 || (size_t)n*16>(size_t)( r.m_pEnd-r.m_p ) )
			r.m_bError = true;
		for ( int i=0; i<n && !r.m_bError; ++i )
		{
			r.GetInt(); r.GetInt(); r.GetInt(); r.GetInt();
		}
		if ( n>0 )
			q.m_sError = "filter trees are not supported";
	}
	if ( r.m_bError )
	{
		sFatal = "invalid or truncated request";
		return false;
	}
	return true;
}

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

} // namespace
} // namespace mgpu

using namespace mgpu;

struct mgpu_api
{
	mgpu_index *				m_pIndex = nullptr;
	mgpu_sharded *				m_pSharded = nullptr;	///< set instead of m_pIndex: rowid-range shards, results carry global rowids
	IndexHeader_t				m_tHdr;
	std::vector<std::unique_ptr<Mapped_t>> m_dSpa;		///< the shards' attribute rows, in shard order
	std::vector<uint64_t>		m_dBase;				///< global rowid of every shard's row 0, plus the total at the end
	int							m_iStride = 0;

	bool HasIndex () const		{ return m_pIndex || m_pSharded; }
	/// the attribute row of a (global) rowid
	const DWORD * Row ( uint32_t uRowid ) const
	{
		size_t s = 0;
		while ( s+1<m_dSpa.size() && uRowid>=m_dBase[s+1] )
			++s;
		return (const DWORD *)m_dSpa[s]->m_p + (size_t)( uRowid-m_dBase[s] )*m_iStride;
	}
	mgpu_parser_settings		m_tTok {};
	std::vector<std::string>	m_dFieldNames, m_dStopwords;
	std::vector<const char *>	m_dFieldPtrs, m_dStopPtrs;
	std::vector<BYTE>			m_dReply;
	std::string					m_sDescribe, m_sError;

	int AttrIndex ( const std::string & sName ) const
	{
		for ( size_t i=0; i<m_tHdr.m_dAttrs.size(); ++i )
			if ( !strcasecmp ( m_tHdr.m_dAttrs[i].m_sName.c_str(), sName.c_str() ) )
				return (int)i;
		return -1;
	}
	int FieldIndex ( const std::string & sName ) const
	{
		for ( size_t i=0; i<m_tHdr.m_dFields.size(); ++i )
			if ( !strcasecmp ( m_tHdr.m_dFields[i].m_sName.c_str(), sName.c_str() ) )
				return (int)i;
		return -1;
	}
};

namespace mgpu
{
namespace
{

/// the query as a SphinxQL-like line (what query_log_format=sphinxql writes for an API query, LogQuerySphinxql in src/searchd.cpp)
static std::string Describe ( const ApiQuery_t & q )
{
	static const char * dModes[] = { "all", "any", "phrase", "boolean", "extended", "fullscan", "extended2" };
	static const char * dSorts[] = { "relevance", "attr_desc", "attr_asc", "time_segments", "extended", "expr" };
	std::string s = "SELECT " + q.m_sSelect + " FROM " + q.m_sIndexes + " WHERE MATCH('" + q.m_sRawQuery + "')";
	for ( const auto & f : q.m_dFilters )
	{
		s += " AND " + f.m_sAttr;
		if ( f.m_eType==SPH_FILTER_RANGE )
			s += std::string ( f.m_bExclude ? " NOT" : "" ) + " BETWEEN " + std::to_string ( f.m_iMin ) + " AND " + std::to_string ( f.m_iMax );
		else if ( f.m_eType==SPH_FILTER_VALUES )
		{
			s += f.m_bExclude ? " NOT IN (" : " IN (";
			for ( size_t i=0; i<f.m_dValues.size(); ++i )
				s += ( i ? "," : "" ) + std::to_string ( f.m_dValues[i] );
			s += ")";
		} else
			s += " <filter type " + std::to_string ( f.m_eType ) + ">";
	}
	if ( !q.m_sGroupBy.empty() )
		s += " GROUP BY " + q.m_sGroupBy;
	s += " ORDER BY " + std::string ( q.m_eSort>=0 && q.m_eSort<=5 ? dSorts[q.m_eSort] : "?" ) + ( q.m_sSortBy.empty() ? "" : "(" + q.m_sSortBy + ")" );
	s += " LIMIT " + std::to_string ( q.m_iOffset ) + "," + std::to_string ( q.m_iLimit );
	s += " OPTION mode=" + std::string ( q.m_eMode>=0 && q.m_eMode<=6 ? dModes[q.m_eMode] : "?" ) + ", ranker=" + std::to_string ( q.m_eRanker )
		+ ", max_matches=" + std::to_string ( q.m_iMaxMatches );
	if ( !q.m_dWeights.empty() )
	{
		s += ", weights=(";
		for ( size_t i=0; i<q.m_dWeights.size(); ++i )
			s += ( i ? "," : "" ) + std::to_string ( q.m_dWeights[i] );
		s += ")";
	}
	if ( !q.m_dFieldWeights.empty() )
	{
		s += ", field_weights=(";
		for ( size_t i=0; i<q.m_dFieldWeights.size(); ++i )
			s += ( i ? "," : "" ) + q.m_dFieldWeights[i].first + "=" + std::to_string ( q.m_dFieldWeights[i].second );
		s += ")";
	}
	if ( q.m_uFlags & QFLAG_PLAIN_IDF )
		s += ", idf=plain";
	if ( q.m_iCutoff )
		s += ", cutoff=" + std::to_string ( q.m_iCutoff );
	if ( !q.m_sComment.empty() )
		s += " /* " + q.m_sComment + " */";
	return s;
}

/// one clause of the extended sort mode: "@weight desc, price asc, @id asc" (sphParseSortClause, src/sortsetup.cpp)
static bool ParseSortClause ( const mgpu_api & A, const std::string & sClause, std::vector<mgpu_sortkey> & dKeys, std::string & sError )
{
	size_t i = 0;
	const size_t n = sClause.size();
	while ( i<n )
	{
		while ( i<n && ( isspace ( (unsigned char)sClause[i] ) || sClause[i]==',' ) )
			++i;
		if ( i>=n )
			break;
		size_t j = i;
		while ( j<n && !isspace ( (unsigned char)sClause[j] ) && sClause[j]!=',' )
			++j;
		std::string sKey = sClause.substr ( i, j-i );
		i = j;
		while ( i<n && isspace ( (unsigned char)sClause[i] ) )
			++i;
		j = i;
		while ( j<n && isalpha ( (unsigned char)sClause[j] ) )
			++j;
		std::string sDir = ToLower ( sClause.substr ( i, j-i ) );
		bool bDesc = false;
		if ( sDir=="desc" )
			bDesc = true;
		else if ( sDir!="asc" )
		{
			sError = "invalid sorting order '" + sDir + "'";		// the reference requires asc / desc after every key
			return false;
		}
		i = j;
		mgpu_sortkey k {};
		k.desc = bDesc;
		if ( sKey=="@weight" || sKey=="@rank" || sKey=="@relevance" || sKey=="weight()" )
			k.kind = MGPU_KEYPART_WEIGHT;
		else
		{
			if ( sKey=="@id" )
				sKey = "id";
			const int iAttr = A.AttrIndex ( sKey );
			if ( iAttr<0 )
			{
				sError = "sort-by attribute '" + sKey + "' not found";
				return false;
			}
			k.kind = A.m_tHdr.m_dAttrs[iAttr].m_eType==SPH_ATTR_FLOAT ? MGPU_KEYPART_FLOAT : MGPU_KEYPART_INT;
			k.attr = iAttr;
		}
		dKeys.push_back ( k );
		if ( dKeys.size()>5 )
		{
			sError = "too many sort-by attributes; maximum count is 5";
			return false;
		}
	}
	if ( dKeys.empty() )
	{
		sError = "empty sort-by clause";
		return false;
	}
	return true;
}

struct Prepared_t
{
	std::string					m_sError;
	mgpu_parsed *				m_pParsed = nullptr;
	mgpu_query					m_tQuery {};
	std::vector<mgpu_sortkey>	m_dSortKeys;
	std::vector<mgpu_filter>	m_dFilters;
	std::vector<int32_t>		m_dFieldWeights;
	std::vector<uint32_t>		m_dRowid;
	std::vector<int32_t>		m_dWeight;
	std::vector<int64_t>		m_dDocid;
	std::vector<mgpu_wordstat>	m_dWordStats;
	mgpu_result					m_tResult {};
	int							m_iBatchSlot = -1;
	~Prepared_t() { mgpu_parsed_free ( m_pParsed ); }
};

/// CSphQuery -> mgpu_query: what sphCreateQueue / BindWeights / CreateFilter decide from the request
static void Prepare ( const mgpu_api & A, const ApiQuery_t & q, Prepared_t & P )
{
	if ( !q.m_sError.empty() ) { P.m_sError = q.m_sError; return; }
	if ( !q.m_sGroupBy.empty() ) { P.m_sError = "group-by is not supported"; return; }
	if ( q.m_eRanker==SPH_RANK_EXPR || q.m_eRanker==SPH_RANK_EXPORT || q.m_eRanker<0 || q.m_eRanker>7 ) { P.m_sError = "expression rankers are not supported"; return; }
	if ( q.m_bGeoAnchor ) { P.m_sError = "geo anchors are not supported"; return; }
	if ( q.m_sSelect!="*" ) { P.m_sError = "select lists other than * are not supported"; return; }
	if ( q.m_bHasOuter ) { P.m_sError = "outer order is not supported"; return; }
	if ( q.m_bTokenFilter ) { P.m_sError = "query token filters are not supported"; return; }
	if ( q.m_iCutoff>0 ) { P.m_sError = "cutoff is not supported"; return; }
	if ( q.m_eMode==5 ) { P.m_sError = "fullscan mode is not supported"; return; }
	if ( q.m_eMode<0 || q.m_eMode>6 ) { P.m_sError = "invalid match mode " + std::to_string ( q.m_eMode ); return; }
	if ( q.m_iMaxMatches<1 || q.m_iMaxMatches>65536 ) { P.m_sError = "max_matches out of bounds (1..65536)"; return; }
	if ( q.m_iOffset<0 || q.m_iLimit<0 || q.m_iOffset>=q.m_iMaxMatches ) { P.m_sError = "offset out of bounds (offset=" + std::to_string ( q.m_iOffset ) + ", max_matches=" + std::to_string ( q.m_iMaxMatches ) + ")"; return; }

	// the tree
	mgpu_parser_settings tTok = A.m_tTok;
	static const int dModeMap[] = { MGPU_MATCH_ALL, MGPU_MATCH_ANY, MGPU_MATCH_PHRASE, MGPU_MATCH_BOOLEAN, MGPU_MATCH_EXTENDED, -1, MGPU_MATCH_EXTENDED };
	tTok.match_mode = dModeMap[q.m_eMode];
	int iRes = mgpu_parse_query ( &tTok, q.m_sRawQuery.c_str(), &P.m_pParsed );
	if ( iRes!=MGPU_OK )
	{
		P.m_sError = std::string ( "index " ) + q.m_sIndexes + ": query error: " + ( P.m_pParsed ? mgpu_parsed_error ( P.m_pParsed ) : "parser failed" );
		return;
	}
	mgpu_query & t = P.m_tQuery;
	t.ranker = q.m_eRanker;		// ESphRankMode values 0..7 are the MGPU_RANK_* values
	mgpu_parsed_fill ( P.m_pParsed, &t );	// (a legacy match mode overrides the ranker, PrepareQueryEmulation)
	t.max_matches = q.m_iMaxMatches;
	t.plain_idf = ( q.m_uFlags & QFLAG_PLAIN_IDF )!=0;

	// weights: named ones win over the positional list (BindWeights, src/sphinx.cpp:13903-13947)
	const int nFields = (int)A.m_tHdr.m_dFields.size();
	if ( !q.m_dFieldWeights.empty() || !q.m_dWeights.empty() )
	{
		P.m_dFieldWeights.assign ( nFields, 1 );
		if ( !q.m_dFieldWeights.empty() )
		{
			for ( const auto & w : q.m_dFieldWeights )
			{
				const int iField = A.FieldIndex ( w.first );
				if ( iField>=0 )
					P.m_dFieldWeights[iField] = w.second;
			}
		} else
			for ( int i=0; i<std::min<int> ( nFields, (int)q.m_dWeights.size() ); ++i )
				P.m_dFieldWeights[i] = (int)q.m_dWeights[i];
		t.field_weights = P.m_dFieldWeights.data();
		t.n_field_weights = nFields;
	}

	// sorting
	switch ( q.m_eSort )
	{
	case SPH_SORT_RELEVANCE:
		break;
	case SPH_SORT_ATTR_DESC:
	case SPH_SORT_ATTR_ASC:
		{
			// MatchAttrLt_fn / MatchAttrGt_fn (src/sphinxsort.cpp:4560-4600): the attribute, then weight descending, then rowid ascending
			const int iAttr = A.AttrIndex ( q.m_sSortBy=="@id" ? "id" : q.m_sSortBy );
			if ( iAttr<0 ) { P.m_sError = "sort-by attribute '" + q.m_sSortBy + "' not found"; return; }
			mgpu_sortkey k {};
			k.kind = A.m_tHdr.m_dAttrs[iAttr].m_eType==SPH_ATTR_FLOAT ? MGPU_KEYPART_FLOAT : MGPU_KEYPART_INT;
			k.attr = iAttr;
			k.desc = q.m_eSort==SPH_SORT_ATTR_DESC;
			P.m_dSortKeys.push_back ( k );
			mgpu_sortkey w {};
			w.kind = MGPU_KEYPART_WEIGHT;
			w.desc = 1;
			P.m_dSortKeys.push_back ( w );
		}
		break;
	case SPH_SORT_EXTENDED:
		if ( !ParseSortClause ( A, q.m_sSortBy, P.m_dSortKeys, P.m_sError ) )
			return;
		break;
	default:
		P.m_sError = "time-segment and expression sort modes are not supported";
		return;
	}
	if ( !P.m_dSortKeys.empty() )
	{
		t.sort_keys = P.m_dSortKeys.data();
		t.n_sort_keys = (int)P.m_dSortKeys.size();
	}

	// filters
	for ( const auto & f : q.m_dFilters )
	{
		if ( !f.m_sUnsupported.empty() ) { P.m_sError = f.m_sUnsupported; return; }
		const int iAttr = A.AttrIndex ( f.m_sAttr=="@id" ? "id" : f.m_sAttr );
		if ( iAttr<0 ) { P.m_sError = "no such filter attribute '" + f.m_sAttr + "'"; return; }
		mgpu_filter m {};
		m.attr = iAttr;
		m.exclude = f.m_bExclude;
		if ( f.m_eType==SPH_FILTER_RANGE )
		{
			m.kind = MGPU_FILTER_RANGE;
			m.min_value = f.m_iMin;
			m.max_value = f.m_iMax;
		} else
		{
			m.kind = MGPU_FILTER_VALUES;
			m.values = f.m_dValues.data();
			m.n_values = (int)f.m_dValues.size();
		}
		P.m_dFilters.push_back ( m );
	}
	if ( !P.m_dFilters.empty() )
	{
		t.filters = P.m_dFilters.data();
		t.n_filters = (int)P.m_dFilters.size();
	}

	P.m_dRowid.resize ( q.m_iMaxMatches );
	P.m_dWeight.resize ( q.m_iMaxMatches );
	P.m_dDocid.resize ( q.m_iMaxMatches );
	P.m_dWordStats.resize ( std::max ( t.n_words, 1 ) );
	P.m_tResult.rowid = P.m_dRowid.data();
	P.m_tResult.weight = P.m_dWeight.data();
	P.m_tResult.docid = P.m_dDocid.data();
	P.m_tResult.word_stats = P.m_dWordStats.data();
}

/// SendResult, src/searchd.cpp:3398-3510 (client mode)
static void SendResult ( const mgpu_api & A, const ApiQuery_t & q, const Prepared_t & P, int iQueryMsec, NetWriter_t & tOut )
{
	if ( !P.m_sError.empty() )
	{
		tOut.SendInt ( SEARCHD_ERROR );
		tOut.SendString ( P.m_sError );
		return;
	}
	const char * sWarning = mgpu_parsed_warning ( P.m_pParsed );
	if ( sWarning && *sWarning )
	{
		tOut.SendDword ( SEARCHD_WARNING );
		tOut.SendString ( sWarning );
	} else
		tOut.SendDword ( SEARCHD_OK );

	// SendSchema: fields, then every attribute but the document id (sphGetAttrsToSend)
	tOut.SendInt ( (int)A.m_tHdr.m_dFields.size() );
	for ( const auto & f : A.m_tHdr.m_dFields )
		tOut.SendString ( f.m_sName );
	tOut.SendInt ( (int)A.m_tHdr.m_dAttrs.size()-1 );
	for ( size_t i=1; i<A.m_tHdr.m_dAttrs.size(); ++i )
	{
		tOut.SendString ( A.m_tHdr.m_dAttrs[i].m_sName );
		tOut.SendDword ( A.m_tHdr.m_dAttrs[i].m_eType );
	}

	const mgpu_result & r = P.m_tResult;
	const int iFrom = std::min ( q.m_iOffset, r.n_matches );
	const int iCount = std::max ( 0, std::min ( q.m_iLimit, r.n_matches-iFrom ) );
	tOut.SendInt ( iCount );
	tOut.SendInt ( 1 );		// 64-bit ids
	for ( int i=iFrom; i<iFrom+iCount; ++i )
	{
		tOut.SendUint64 ( (uint64_t)r.docid[i] );
		tOut.SendInt ( r.weight[i] );
		const DWORD * pRow = A.Row ( r.rowid[i] );
		for ( size_t a=1; a<A.m_tHdr.m_dAttrs.size(); ++a )
		{
			const SchemaAttr_t & t = A.m_tHdr.m_dAttrs[a];
			const DWORD * p = pRow + t.m_iBitOffset/32;
			if ( t.m_iBitCount==64 )
				tOut.SendUint64 ( (uint64_t)p[0] | ( (uint64_t)p[1]<<32 ) );
			else if ( t.m_iBitCount==32 )
				tOut.SendDword ( p[0] );	// integers, timestamps and floats travel as their 32 bits (SendAttribute)
			else
				tOut.SendDword ( ( p[0]>>( t.m_iBitOffset & 31 ) ) & ( ( 1u<<t.m_iBitCount )-1 ) );
		}
	}
	tOut.SendInt ( r.n_matches );			// "total": what the sorter kept
	tOut.SendAsDword ( r.total_found );
	tOut.SendInt ( std::max ( iQueryMsec, 0 ) );

	// per-keyword statistics, sorted by keyword, one line per distinct keyword (MakeSortedWordStat)
	std::map<std::string,mgpu_wordstat> hWords;
	for ( int w=0; w<P.m_tQuery.n_words; ++w )
		if ( P.m_tQuery.words[w].word )
			hWords[P.m_tQuery.words[w].word] = r.word_stats[w];
	tOut.SendInt ( (int)hWords.size() );
	for ( const auto & kv : hWords )
	{
		tOut.SendString ( kv.first );
		tOut.SendAsDword ( kv.second.docs );
		tOut.SendAsDword ( kv.second.hits );
	}
}

static void SendErrorReply ( NetWriter_t & tOut, const std::string & sError )
{
	// SendErrorReply (src/searchd.cpp): status SEARCHD_ERROR, version 0, the message
	tOut.m_d.clear();
	tOut.SendWord ( SEARCHD_ERROR );
	tOut.SendWord ( 0 );
	tOut.SendDword ( (uint32_t)sError.size()+4 );
	tOut.SendString ( sError );
}

} // namespace
} // namespace mgpu

extern "C"
{

static int ApiCreateImpl ( mgpu_index * idx, mgpu_sharded * sh, const char * const * path_prefixes, int n_prefixes, const mgpu_parser_settings * tokenizer, mgpu_api ** out );

int mgpu_api_create ( mgpu_index * idx, const char * path_prefix, const mgpu_parser_settings * tokenizer, mgpu_api ** out )
{
	if ( !path_prefix || !out )
		return MGPU_E_BAD_QUERY;
	*out = nullptr;
	try
	{
		return ApiCreateImpl ( idx, nullptr, &path_prefix, 1, tokenizer, out );
	} catch ( ... )
	{
		return MGPU_E_NOMEM;
	}
}

static int ApiCreateImpl ( mgpu_index * idx, mgpu_sharded * sh, const char * const * path_prefixes, int n_prefixes, const mgpu_parser_settings * tokenizer, mgpu_api ** out )
{
	std::unique_ptr<mgpu_api> p ( new mgpu_api );
	p->m_pIndex = idx;
	p->m_pSharded = sh;
	std::string sError;
	uint64_t uBase = 0;
	for ( int s=0; s<n_prefixes; ++s )
	{
		if ( !path_prefixes[s] )
			return MGPU_E_BAD_QUERY;
		const std::string sPrefix ( path_prefixes[s] );
		Mapped_t tSph;
		std::unique_ptr<Mapped_t> pSpa ( new Mapped_t );
		if ( !tSph.Map ( sPrefix+".sph" ) || !pSpa->Map ( sPrefix+".spa" ) )
			return MGPU_E_IO;
		IndexHeader_t tHdr;
		if ( !ReadHeader ( tSph.m_p, tSph.m_iLen, tHdr, sError ) || tHdr.m_dAttrs.empty() || tHdr.m_iDocinfo<0 )
			return MGPU_E_FORMAT;
		const int iStride = tHdr.RowStride();
		if ( (int64_t)pSpa->m_iLen/4/std::max ( iStride, 1 )<tHdr.m_iDocinfo )
			return MGPU_E_FORMAT;
		if ( s==0 )
		{
			p->m_tHdr = tHdr;
			p->m_iStride = iStride;
		} else
		{
			// shards of one index share its schema
			bool bSame = tHdr.m_dAttrs.size()==p->m_tHdr.m_dAttrs.size() && tHdr.m_dFields.size()==p->m_tHdr.m_dFields.size() && iStride==p->m_iStride;
			for ( size_t a=0; bSame && a<tHdr.m_dAttrs.size(); ++a )
				bSame = tHdr.m_dAttrs[a].m_sName==p->m_tHdr.m_dAttrs[a].m_sName && tHdr.m_dAttrs[a].m_iBitOffset==p->m_tHdr.m_dAttrs[a].m_iBitOffset
					&& tHdr.m_dAttrs[a].m_iBitCount==p->m_tHdr.m_dAttrs[a].m_iBitCount;
			if ( !bSame )
				return MGPU_E_FORMAT;
		}
		p->m_dSpa.push_back ( std::move ( pSpa ) );
		p->m_dBase.push_back ( uBase );
		uBase += (uint64_t)tHdr.m_iDocinfo;
	}
	p->m_dBase.push_back ( uBase );

	// the tokenizer settings are copied; the field names are always the index's own
	if ( tokenizer )
	{
		p->m_tTok = *tokenizer;
		for ( int i=0; i<tokenizer->n_stopwords && tokenizer->stopwords; ++i )
			if ( tokenizer->stopwords[i] )
				p->m_dStopwords.push_back ( tokenizer->stopwords[i] );
	} else
	{
		p->m_tTok.min_word_len = (int)p->m_tHdr.m_iMinWordLen;
		p->m_tTok.overshort_step = (int)p->m_tHdr.m_iOvershortStep;
		p->m_tTok.stopword_step = (int)p->m_tHdr.m_iStopwordStep;
		p->m_tTok.ngram_cjk = 1;
	}
	for ( const auto & f : p->m_tHdr.m_dFields )
		p->m_dFieldNames.push_back ( f.m_sName );
	for ( const auto & s : p->m_dFieldNames )
		p->m_dFieldPtrs.push_back ( s.c_str() );
	for ( const auto & s : p->m_dStopwords )
		p->m_dStopPtrs.push_back ( s.c_str() );
	p->m_tTok.n_fields = (int)p->m_dFieldPtrs.size();
	p->m_tTok.field_names = p->m_dFieldPtrs.data();
	p->m_tTok.n_stopwords = (int)p->m_dStopPtrs.size();
	p->m_tTok.stopwords = p->m_dStopPtrs.empty() ? nullptr : p->m_dStopPtrs.data();
	*out = p.release();
	return MGPU_OK;
}

int mgpu_api_create_sharded ( mgpu_sharded * sh, const char * const * path_prefixes, int n_shards, const mgpu_parser_settings * tokenizer, mgpu_api ** out )
{
	if ( !path_prefixes || n_shards<1 || !out )
		return MGPU_E_BAD_QUERY;
	*out = nullptr;
	try
	{
		return ApiCreateImpl ( nullptr, sh, path_prefixes, n_shards, tokenizer, out );
	} catch ( ... )
	{
		return MGPU_E_NOMEM;
	}
}

void mgpu_api_free ( mgpu_api * api )
{
	delete api;
}

const char * mgpu_api_describe_last ( const mgpu_api * api )
{
	return api ? api->m_sDescribe.c_str() : "";
}

static int ApiHandleImpl ( mgpu_api * api, const void * request, size_t request_len, const void ** reply, size_t * reply_len );

int mgpu_api_handle ( mgpu_api * api, const void * request, size_t request_len, const void ** reply, size_t * reply_len )
{
	// no exception crosses the ABI (a packet can ask for allocations up to its own size; bad_alloc is the one thing left to catch)
	try
	{
		return ApiHandleImpl ( api, request, request_len, reply, reply_len );
	} catch ( ... )
	{
		return MGPU_E_NOMEM;
	}
}

static int ApiHandleImpl ( mgpu_api * api, const void * request, size_t request_len, const void ** reply, size_t * reply_len )
{
	if ( !api || !reply || !reply_len || ( request_len && !request ) )
		return MGPU_E_BAD_QUERY;
	mgpu_api & A = *api;
	A.m_sDescribe.clear();
	NetWriter_t tOut;
	auto fnDone = [&] ()
	{
		A.m_dReply.swap ( tOut.m_d );
		*reply = A.m_dReply.data();
		*reply_len = A.m_dReply.size();
		return MGPU_OK;
	};

	// the packet: command, command version, body length (src/netreceive_api.cpp), then HandleCommandSearch (src/searchd.cpp:6932-7000)
	NetReader_t r ( request, request_len );
	const uint16_t uCommand = r.GetWord();
	const uint16_t uVer = r.GetWord();
	const uint32_t uLen = r.GetDword();
	if ( r.m_bError || uLen!=(size_t)( r.m_pEnd-r.m_p ) )
	{
		SendErrorReply ( tOut, "invalid or truncated request" );
		return fnDone();
	}
	if ( uCommand==SEARCHD_COMMAND_KEYWORDS )
	{
		// HandleCommandKeywords (src/searchd.cpp): the text tokenized as the index would tokenize it, optionally with the dictionary's counts
		char sBuf[160];
		if ( ( uVer>>8 )!=( VER_COMMAND_KEYWORDS>>8 ) || uVer>VER_COMMAND_KEYWORDS )
		{
			if ( ( uVer>>8 )!=( VER_COMMAND_KEYWORDS>>8 ) )
				snprintf ( sBuf, sizeof(sBuf), "major command version mismatch (expected v.%d.x, got v.%d.%d)", VER_COMMAND_KEYWORDS>>8, uVer>>8, uVer & 255 );
			else
				snprintf ( sBuf, sizeof(sBuf), "client version is higher than daemon version (client is v.%d.%d, daemon is v.%d.%d)", uVer>>8, uVer & 255, VER_COMMAND_KEYWORDS>>8, VER_COMMAND_KEYWORDS & 255 );
			SendErrorReply ( tOut, sBuf );
			return fnDone();
		}
		const std::string sQuery = r.GetString();
		const std::string sIndex = r.GetString();
		const bool bStats = r.GetInt()!=0;
		if ( uVer>=0x101 )
		{
			r.GetInt(); r.GetInt(); r.GetInt(); r.GetInt();		// fold lemmas / blended / wildcards, expansion limit: nothing to fold here
		}
		if ( r.m_bError || r.m_p!=r.m_pEnd )
		{
			SendErrorReply ( tOut, "invalid or truncated request" );
			return fnDone();
		}
		if ( bStats && !A.HasIndex() )
		{
			SendErrorReply ( tOut, "no index is attached to this responder" );
			return fnDone();
		}
		A.m_sDescribe = "CALL KEYWORDS('" + sQuery + "', '" + sIndex + "', " + ( bStats ? "1" : "0" ) + ");\n";
		std::vector<std::pair<std::string,int>> dWords;
		TokenizePlain ( A.m_tTok, sQuery.c_str(), dWords );
		tOut.SendWord ( SEARCHD_OK );
		tOut.SendWord ( VER_COMMAND_KEYWORDS );
		tOut.SendDword ( 0 );
		tOut.SendInt ( (int)dWords.size() );
		for ( const auto & w : dWords )
		{
			tOut.SendString ( w.first );	// tokenized
			tOut.SendString ( w.first );	// normalized: no morphology on this path
			if ( uVer>=0x101 )
				tOut.SendInt ( w.second );
			if ( bStats )
			{
				int64_t iDocs = 0, iHits = 0;
				if ( A.m_pSharded )
					mgpu_sharded_word_stats ( A.m_pSharded, w.first.c_str(), &iDocs, &iHits );
				else
					mgpu_index_word_stats ( A.m_pIndex, w.first.c_str(), &iDocs, &iHits );
				tOut.SendAsDword ( iDocs );
				tOut.SendAsDword ( iHits );
			}
		}
		tOut.PatchDword ( 4, (uint32_t)( tOut.m_d.size()-8 ) );
		return fnDone();
	}
	if ( uCommand!=SEARCHD_COMMAND_SEARCH )
	{
		SendErrorReply ( tOut, "unknown command (code=" + std::to_string ( uCommand ) + ")" );
		return fnDone();
	}
	// CheckCommandVersion (src/searchd.cpp): same major version, client minor not above the daemon's; clients older than 1.29 lack
	// fields this responder reads unconditionally
	{
		char s[160];
		if ( ( uVer>>8 )!=( VER_COMMAND_SEARCH>>8 ) )
		{
			snprintf ( s, sizeof(s), "major command version mismatch (expected v.%d.x, got v.%d.%d)", VER_COMMAND_SEARCH>>8, uVer>>8, uVer & 255 );
			SendErrorReply ( tOut, s );
			return fnDone();
		}
		if ( uVer>VER_COMMAND_SEARCH )
		{
			snprintf ( s, sizeof(s), "client version is higher than daemon version (client is v.%d.%d, daemon is v.%d.%d)", uVer>>8, uVer & 255, VER_COMMAND_SEARCH>>8, VER_COMMAND_SEARCH & 255 );
			SendErrorReply ( tOut, s );
			return fnDone();
		}
		if ( uVer<0x11D )
		{
			snprintf ( s, sizeof(s), "client version v.%d.%d is too old (v.1.29 or newer is served)", uVer>>8, uVer & 255 );
			SendErrorReply ( tOut, s );
			return fnDone();
		}
	}
	const int iMasterVer = r.GetInt();
	const int nQueries = r.GetInt();
	if ( iMasterVer!=0 )
	{
		SendErrorReply ( tOut, "master-agent extensions are not supported" );
		return fnDone();
	}
	if ( nQueries<=0 || nQueries>32 )	// the reference's max_batch_queries default
	{
		SendErrorReply ( tOut, "bad multi-query count " + std::to_string ( nQueries ) + " (must be in 1..32 range)" );
		return fnDone();
	}

	std::vector<ApiQuery_t> dQueries ( nQueries );
	for ( auto & q : dQueries )
	{
		std::string sFatal;
		if ( !ParseQuery ( r, q, uVer, sFatal ) )
		{
			SendErrorReply ( tOut, sFatal );
			return fnDone();
		}
		A.m_sDescribe += Describe ( q ) + ";\n";
	}
	if ( r.m_p!=r.m_pEnd )
	{
		SendErrorReply ( tOut, "invalid or truncated request" );
		return fnDone();
	}

	// run what can run as ONE batch
	const auto tStart = std::chrono::steady_clock::now();
	std::vector<std::unique_ptr<Prepared_t>> dPrepared;
	std::vector<mgpu_query> dBatch;
	std::vector<mgpu_result> dResults;
	for ( auto & q : dQueries )
	{
		dPrepared.emplace_back ( new Prepared_t );
		Prepared_t & P = *dPrepared.back();
		Prepare ( A, q, P );
		if ( P.m_sError.empty() && !A.HasIndex() )
			P.m_sError = "no index is attached to this responder";
		if ( P.m_sError.empty() )
		{
			P.m_iBatchSlot = (int)dBatch.size();
			dBatch.push_back ( P.m_tQuery );
			dResults.push_back ( P.m_tResult );
		}
	}
	if ( !dBatch.empty() )
	{
		const int iRes = A.m_pSharded ? mgpu_sharded_search_batch ( A.m_pSharded, dBatch.data(), (int)dBatch.size(), dResults.data() )
			: mgpu_search_batch ( A.m_pIndex, dBatch.data(), (int)dBatch.size(), dResults.data() );
		const char * sLast = A.m_pSharded ? mgpu_sharded_last_error ( A.m_pSharded ) : mgpu_last_error ( A.m_pIndex );
		for ( auto & pP : dPrepared )
			if ( pP->m_iBatchSlot>=0 )
			{
				pP->m_tResult = dResults[pP->m_iBatchSlot];
				const int iStatus = iRes!=MGPU_OK ? iRes : pP->m_tResult.status;
				if ( iStatus==MGPU_E_UNSUPPORTED )
					pP->m_sError = "query uses an operator or option the GPU path does not implement";
				else if ( iStatus!=MGPU_OK )
					pP->m_sError = std::string ( "search failed: " ) + ( sLast ? sLast : "" ) + " (code " + std::to_string ( iStatus ) + ")";
			}
	}
	const int iMsec = (int)std::chrono::duration_cast<std::chrono::milliseconds> ( std::chrono::steady_clock::now()-tStart ).count();

	// APIAnswer: status, version, length; then one result per query
	tOut.SendWord ( SEARCHD_OK );
	tOut.SendWord ( VER_COMMAND_SEARCH );
	tOut.SendDword ( 0 );
	for ( int i=0; i<nQueries; ++i )
		SendResult ( A, dQueries[i], *dPrepared[i], iMsec, tOut );
	tOut.PatchDword ( 4, (uint32_t)( tOut.m_d.size()-8 ) );
	return fnDone();
}

} // extern "C"
