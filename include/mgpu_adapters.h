// mgpu_adapters.h -- C++ host-side mirror of the reference's operator surface over the C ABI of mgpu.h.
//
// Header-only, in the reference's idiom (_c/_t/_i suffixes, m_ members). It lets a caller written against Manticore's
// types keep its shape:
//   XQNode_t / XQKeyword_t / XQLimitSpec_t   (src/sphinxquery.h:21-39, 65-129, 134-280)  ->  GpuXQNode_t / GpuXQKeyword_t
//   CSphQuery subset                          (src/sphinx.h:2586-2691)                    ->  GpuQuery_t
//   CSphMatch + ISphMatchSorter::Push/Flatten (src/sphinx.h:1104, src/sphinxsort.h:39-133) ->  GpuMatch_t / GpuMatchSorter_i / GpuMatchQueue_c
//   CSphIndex::MultiQuery / MultiQueryEx      (src/sphinx.h:3169-3173)                    ->  GpuIndex_c::MultiQuery / MultiQueryEx
//   CSphQueryResultMeta (m_iTotalMatches, AddStat, m_sError)  (src/sphinx.h:2709+)        ->  GpuQueryResultMeta_t
// Semantics follow the reference: MultiQuery returns bool + m_sError; results are Push()ed into caller-owned sorters
// (order-insensitive: the comparator is a total order with rowid as the last tie-break, src/sphinxsort.cpp:4534-4548).
#pragma once

#include "mgpu.h"

#include <algorithm>
#include <memory>
#include <string>
#include <vector>

/// XQKeyword_t (src/sphinxquery.h:21-39)
struct GpuXQKeyword_t
{
	std::string	m_sWord;
	int			m_iAtomPos = -1;
	float		m_fBoost = 1.0f;
	bool		m_bFieldStart = false;
	bool		m_bFieldEnd = false;
	bool		m_bExcluded = false;
	bool		m_bExpanded = false;
};

/// XQNode_t (src/sphinxquery.h:134-280) with the XQLimitSpec_t members the GPU path understands
struct GpuXQNode_t
{
	int			m_eOp = MGPU_OP_AND;			///< XQOperator_e
	int			m_iOpArg = 0;
	uint32_t	m_uFieldMask = 0xFFFFFFFFu;		///< m_dSpec.m_dFieldMask (fields 0..31)
	int			m_iFieldMaxPos = 0;				///< m_dSpec.m_iFieldMaxPos
	bool		m_bNotWeighted = false;
	std::vector<std::unique_ptr<GpuXQNode_t>>	m_dChildren;
	std::vector<GpuXQKeyword_t>					m_dWords;

	static std::unique_ptr<GpuXQNode_t> Keyword ( const char * sWord, int iAtomPos, uint32_t uFieldMask=0xFFFFFFFFu )
	{
		std::unique_ptr<GpuXQNode_t> p ( new GpuXQNode_t );
		p->m_uFieldMask = uFieldMask;
		GpuXQKeyword_t tWord;
		tWord.m_sWord = sWord;
		tWord.m_iAtomPos = iAtomPos;
		p->m_dWords.push_back ( tWord );
		return p;
	}
	static std::unique_ptr<GpuXQNode_t> Op ( int eOp, std::vector<std::unique_ptr<GpuXQNode_t>> dChildren, int iOpArg=0 )
	{
		std::unique_ptr<GpuXQNode_t> p ( new GpuXQNode_t );
		p->m_eOp = eOp;
		p->m_iOpArg = iOpArg;
		p->m_dChildren = std::move ( dChildren );
		return p;
	}
};

/// the CSphQuery fields the hot path reads (src/sphinx.h:2586-2691)
struct GpuQuery_t
{
	std::unique_ptr<GpuXQNode_t>	m_pRoot;			///< XQQuery_t::m_pRoot; null = parse m_sQuery
	std::string		m_sQuery;							///< CSphQuery::m_sQuery: the query text, parsed by mgpu_parse_query (sphParseExtendedQuery + transforms)
	int				m_eMode = MGPU_MATCH_EXTENDED;		///< CSphQuery::m_eMode (legacy modes rewrite the text and pick the ranker, PrepareQueryEmulation)
	int				m_iMinWordLen = 1;					///< the index's tokenizer settings the parser needs
	std::vector<std::string> m_dStopwords;
	int				m_eRanker = MGPU_RANK_PROXIMITY_BM25;	///< m_eRanker
	std::vector<int> m_dFieldWeights;					///< bound weights (CSphQueryContext::m_dWeights)
	std::vector<mgpu_sortkey> m_dSortKeys;				///< CSphMatchComparatorState
	std::vector<mgpu_filter> m_dFilters;
	int				m_iMaxMatches = 1000;				///< DEFAULT_MAX_MATCHES, src/sphinx.h:2583
	int				m_iIndexWeight = 1;
	bool			m_bPlainIDF = false;
	bool			m_bNormalizedTFIDF = true;
	int64_t			m_iTotalDocs = 0;					///< CSphMultiQueryArgs::m_iTotalDocs
};

/// CSphMatch as the sorters see it (src/sphinx.h:1104)
struct GpuMatch_t
{
	uint32_t	m_tRowID = 0;
	int			m_iWeight = 0;
	int64_t		m_iDocID = 0;		///< the `id` attribute of the row
	int64_t		m_iSortAttr = 0;	///< first integer sort key, if any
};

/// per-keyword statistics + totals + error text (CSphQueryResultMeta)
struct GpuQueryResultMeta_t
{
	struct WordStat_t { std::string m_sWord; int64_t m_iDocs; int64_t m_iHits; };
	std::vector<WordStat_t>	m_dWordStats;		///< AddStat order = query-pos order
	int64_t		m_iTotalMatches = 0;		///< total_found
	std::string	m_sError;
	std::string	m_sWarning;
	std::string	m_sParseError;		///< XQQuery_t::m_sParseError (text queries)
	std::string	m_sParseWarning;	///< XQQuery_t::m_sParseWarning
};

/// ISphMatchSorter (src/sphinxsort.h:39-133), the subset MatchExtended and the result merge use
class GpuMatchSorter_i
{
public:
	virtual			~GpuMatchSorter_i() {}
	virtual bool	Push ( const GpuMatch_t & tEntry ) = 0;
	virtual int		GetLength() const = 0;
	virtual int64_t	GetTotalCount() const = 0;
	virtual void	SetTotalCount ( int64_t iTotal ) = 0;	///< the GPU path counts matches on the device
	virtual int		Flatten ( GpuMatch_t * pTo ) = 0;		///< best first (src/sphinxsort.cpp:627-641)
	virtual int		GetMaxMatches() const = 0;
};

/// CSphMatchQueue<MatchRelevanceLt_fn> (src/sphinxsort.cpp:582-812, 4534-4548): weight desc, rowid asc; binary heap, worst at the root
class GpuMatchQueue_c : public GpuMatchSorter_i
{
public:
	explicit GpuMatchQueue_c ( int iSize ) : m_iSize ( iSize ) { m_dData.reserve ( iSize ); }

	static bool IsLess ( const GpuMatch_t & a, const GpuMatch_t & b )
	{
		if ( a.m_iWeight!=b.m_iWeight )
			return a.m_iWeight<b.m_iWeight;
		return a.m_tRowID>b.m_tRowID;
	}
	bool Push ( const GpuMatch_t & tEntry ) override
	{
		++m_iTotal;
		if ( (int)m_dData.size()==m_iSize )
		{
			if ( IsLess ( tEntry, m_dData.front() ) )
				return true;
			std::pop_heap ( m_dData.begin(), m_dData.end(), Better );
			m_dData.pop_back();
		}
		m_dData.push_back ( tEntry );
		std::push_heap ( m_dData.begin(), m_dData.end(), Better );
		return true;
	}
	int GetLength() const override			{ return (int)m_dData.size(); }
	int64_t GetTotalCount() const override	{ return m_iTotal; }
	void SetTotalCount ( int64_t i ) override { m_iTotal = i; }
	int GetMaxMatches() const override		{ return m_iSize; }
	int Flatten ( GpuMatch_t * pTo ) override
	{
		std::vector<GpuMatch_t> d = m_dData;
		std::sort ( d.begin(), d.end(), [] ( const GpuMatch_t & a, const GpuMatch_t & b ) { return IsLess ( b, a ); } );
		std::copy ( d.begin(), d.end(), pTo );
		return (int)d.size();
	}

private:
	static bool Better ( const GpuMatch_t & a, const GpuMatch_t & b )	{ return IsLess ( b, a ); }	// heap top = worst
	std::vector<GpuMatch_t>	m_dData;
	int			m_iSize;
	int64_t		m_iTotal = 0;
};

/// CSphIndex (plain, disk) as the daemon drives it: Prealloc + MultiQuery / MultiQueryEx
class GpuIndex_c
{
public:
	~GpuIndex_c()	{ Dealloc(); }

	/// CSphIndex_VLN::Prealloc (src/sphinx.cpp:13782): loads <prefix>.sph/.spi/.spd/.spp/.spe/.spa onto the device
	bool Prealloc ( const char * sPathPrefix, int iDevice=0, uint32_t uRowidBase=0 )
	{
		Dealloc();
		const int iRes = mgpu_index_open ( sPathPrefix, iDevice, uRowidBase, &m_pIndex );
		if ( iRes!=MGPU_OK )
		{
			m_sLastError = mgpu_last_error ( nullptr );
			m_iLastStatus = iRes;
			return false;
		}
		return true;
	}
	void Dealloc()
	{
		if ( m_pIndex )
			mgpu_index_close ( m_pIndex );
		m_pIndex = nullptr;
	}
	const std::string & GetLastError() const	{ return m_sLastError; }
	int GetLastStatus() const					{ return m_iLastStatus; }
	mgpu_index * Handle() const					{ return m_pIndex; }

	/// CSphIndex::MultiQuery (src/sphinx.h:3169): one query, results Push()ed into the caller's sorter
	bool MultiQuery ( GpuQueryResultMeta_t & tMeta, const GpuQuery_t & tQuery, GpuMatchSorter_i * pSorter )
	{
		return MultiQueryEx ( 1, &tQuery, &tMeta, &pSorter );
	}

	/// CSphIndex::MultiQueryEx (src/sphinx.h:3172): a batch of queries, one sorter and one meta each
	bool MultiQueryEx ( int iQueries, const GpuQuery_t * pQueries, GpuQueryResultMeta_t * pMeta, GpuMatchSorter_i ** ppSorters )
	{
		if ( !m_pIndex )
		{
			for ( int i=0; i<iQueries; ++i )
				pMeta[i].m_sError = "index not preallocated";
			return false;
		}
		std::vector<Flat_t> dFlat ( iQueries );
		std::vector<mgpu_query> dQ ( iQueries );
		std::vector<mgpu_result> dR ( iQueries );
		for ( int i=0; i<iQueries; ++i )
		{
			Flatten ( pQueries[i], dFlat[i], dQ[i] );
			Flat_t & f = dFlat[i];
			if ( !pQueries[i].m_pRoot )
			{
				// the reference's own shape: MultiQuery gets the text and parses it against this index's schema (src/sphinx.cpp:15403-15430)
				std::vector<const char *> dFields, dStops;
				for ( int iField=0; iField<mgpu_index_num_fields ( m_pIndex ); ++iField )
					dFields.push_back ( mgpu_index_field_name ( m_pIndex, iField ) );
				for ( const std::string & s : pQueries[i].m_dStopwords )
					dStops.push_back ( s.c_str() );
				mgpu_parser_settings tTok {};
				tTok.n_fields = (int)dFields.size();	tTok.field_names = dFields.data();
				tTok.n_stopwords = (int)dStops.size();	tTok.stopwords = dStops.empty() ? nullptr : dStops.data();
				tTok.min_word_len = pQueries[i].m_iMinWordLen;
				tTok.overshort_step = 1; tTok.stopword_step = 1; tTok.ngram_cjk = 1;
				tTok.match_mode = pQueries[i].m_eMode;
				const int iParse = mgpu_parse_query ( &tTok, pQueries[i].m_sQuery.c_str(), &f.m_pParsed );
				if ( iParse!=MGPU_OK )
				{
					pMeta[i].m_sParseError = f.m_pParsed ? mgpu_parsed_error ( f.m_pParsed ) : "parser failed";
					dQ[i].n_nodes = 0;	// runs as an empty query; reported below
					dQ[i].root = -1;
				} else
				{
					pMeta[i].m_sParseWarning = mgpu_parsed_warning ( f.m_pParsed );
					mgpu_parsed_fill ( f.m_pParsed, &dQ[i] );
				}
			}
			const int iK = std::max ( 1, ppSorters[i]->GetMaxMatches() );
			dQ[i].max_matches = iK;
			f.m_dRowid.resize ( iK ); f.m_dWeight.resize ( iK ); f.m_dDocid.resize ( iK ); f.m_dSortAttr.resize ( iK );
			f.m_dStats.resize ( (size_t)std::max ( 1, dQ[i].n_words ) );
			dR[i].rowid = f.m_dRowid.data(); dR[i].weight = f.m_dWeight.data(); dR[i].docid = f.m_dDocid.data();
			dR[i].sort_attr = f.m_dSortAttr.data(); dR[i].word_stats = f.m_dStats.data();
		}
		const int iRes = mgpu_search_batch ( m_pIndex, dQ.data(), iQueries, dR.data() );
		bool bOk = ( iRes==MGPU_OK );
		for ( int i=0; i<iQueries; ++i )
		{
			GpuQueryResultMeta_t & tMeta = pMeta[i];
			if ( !tMeta.m_sParseError.empty() )
			{
				tMeta.m_sError = tMeta.m_sParseError;	// XQQuery_t::m_sParseError -> tMeta.m_sError, src/sphinx.cpp:15412
				bOk = false;
				continue;
			}
			if ( iRes!=MGPU_OK || dR[i].status!=MGPU_OK )
			{
				// reference: MultiQuery returns false and sets tMeta.m_sError (src/sphinx.cpp:15690); per query m_iMultiplier=-1
				tMeta.m_sError = dR[i].status==MGPU_E_UNSUPPORTED ? "query not supported by the GPU path" : mgpu_last_error ( m_pIndex );
				bOk = false;
				continue;
			}
			GpuMatch_t tMatch;
			for ( int m=0; m<dR[i].n_matches; ++m )
			{
				tMatch.m_tRowID = dR[i].rowid[m];
				tMatch.m_iWeight = dR[i].weight[m];
				tMatch.m_iDocID = dR[i].docid[m];
				tMatch.m_iSortAttr = dR[i].sort_attr[m];
				ppSorters[i]->Push ( tMatch );
			}
			ppSorters[i]->SetTotalCount ( dR[i].total_found );
			tMeta.m_iTotalMatches = dR[i].total_found;
			tMeta.m_dWordStats.clear();
			for ( int w=0; w<dQ[i].n_words; ++w )	// AddStat, src/sphinxsearch.cpp:4365-4371
				tMeta.m_dWordStats.push_back ( { dQ[i].words[w].word, dFlat[i].m_dStats[w].docs, dFlat[i].m_dStats[w].hits } );
		}
		return bOk;
	}

private:
	struct Flat_t
	{
		std::vector<mgpu_xqnode>	m_dNodes;
		std::vector<int32_t>		m_dChildren;
		std::vector<mgpu_xqkeyword>	m_dWords;
		std::vector<std::string>	m_dWordStrings;
		std::vector<uint32_t>		m_dRowid;
		std::vector<int32_t>		m_dWeight;
		std::vector<int64_t>		m_dDocid, m_dSortAttr;
		std::vector<mgpu_wordstat>	m_dStats;
		mgpu_parsed *				m_pParsed = nullptr;
		Flat_t() = default;
		Flat_t ( const Flat_t & ) = delete;
		~Flat_t() { mgpu_parsed_free ( m_pParsed ); }
	};

	static int FlattenNode ( const GpuXQNode_t * pNode, Flat_t & f )
	{
		const int iMe = (int)f.m_dNodes.size();
		f.m_dNodes.emplace_back();
		mgpu_xqnode t {};
		t.op = pNode->m_eOp;
		t.oparg = pNode->m_iOpArg;
		t.field_mask = pNode->m_uFieldMask;
		t.field_max_pos = pNode->m_iFieldMaxPos;
		t.not_weighted = pNode->m_bNotWeighted;
		t.first_word = (int)f.m_dWords.size();
		t.n_words = (int)pNode->m_dWords.size();
		for ( const GpuXQKeyword_t & w : pNode->m_dWords )
		{
			f.m_dWordStrings.push_back ( w.m_sWord );
			mgpu_xqkeyword k {};
			k.atom_pos = w.m_iAtomPos; k.boost = w.m_fBoost;
			k.field_start = w.m_bFieldStart; k.field_end = w.m_bFieldEnd; k.excluded = w.m_bExcluded; k.expanded = w.m_bExpanded;
			f.m_dWords.push_back ( k );
		}
		std::vector<int> dKids;
		for ( const auto & pChild : pNode->m_dChildren )
			dKids.push_back ( FlattenNode ( pChild.get(), f ) );
		t.first_child = (int)f.m_dChildren.size();
		t.n_children = (int)dKids.size();
		for ( int i : dKids )
			f.m_dChildren.push_back ( i );
		f.m_dNodes[iMe] = t;
		return iMe;
	}

	static void Flatten ( const GpuQuery_t & tQuery, Flat_t & f, mgpu_query & q )
	{
		q = mgpu_query {};
		f.m_dWordStrings.reserve ( 64 );
		q.root = tQuery.m_pRoot ? FlattenNode ( tQuery.m_pRoot.get(), f ) : -1;
		for ( size_t i=0; i<f.m_dWords.size(); ++i )
			f.m_dWords[i].word = f.m_dWordStrings[i].c_str();	// strings are stable from here on
		q.nodes = f.m_dNodes.data();		q.n_nodes = (int)f.m_dNodes.size();
		q.children = f.m_dChildren.data();	q.n_children = (int)f.m_dChildren.size();
		q.words = f.m_dWords.data();		q.n_words = (int)f.m_dWords.size();
		q.ranker = tQuery.m_eRanker;
		q.field_weights = tQuery.m_dFieldWeights.empty() ? nullptr : tQuery.m_dFieldWeights.data();
		q.n_field_weights = (int)tQuery.m_dFieldWeights.size();
		q.sort_keys = tQuery.m_dSortKeys.empty() ? nullptr : tQuery.m_dSortKeys.data();
		q.n_sort_keys = (int)tQuery.m_dSortKeys.size();
		q.filters = tQuery.m_dFilters.empty() ? nullptr : tQuery.m_dFilters.data();
		q.n_filters = (int)tQuery.m_dFilters.size();
		q.max_matches = tQuery.m_iMaxMatches;
		q.index_weight = tQuery.m_iIndexWeight;
		q.plain_idf = tQuery.m_bPlainIDF;
		q.unnormalized_tfidf = !tQuery.m_bNormalizedTFIDF;
		q.total_docs = tQuery.m_iTotalDocs;
	}

	mgpu_index *	m_pIndex = nullptr;
	std::string		m_sLastError;
	int				m_iLastStatus = MGPU_OK;
};
