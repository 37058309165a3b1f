// C++ boundary test in the shape of the reference's own RTN.WeightBoundary (src/gtests/gtests_rtstuff.cpp:244-335):
// build a tiny index, create a sorter, call pIndex->MultiQuery, assert rowid / weight. Plus the field-weight case of
// test/test_322 (negative weights) and a two-query MultiQueryEx batch. Uses only include/mgpu_adapters.h + libmgpu.so.
#include "mgpu_adapters.h"
#include "mgpu_writer.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>

#define CHECK(_c) do { if ( !( _c ) ) { fprintf ( stderr, "CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #_c ); return 1; } } while (0)

struct Corpus_t
{
	std::vector<std::string> m_dFields;
	std::vector<int64_t> m_dIds;
	std::vector<std::vector<std::string>> m_dDocs;	// per doc, per field: text

	bool Build ( const std::string & sPrefix, std::string & sError ) const
	{
		std::map<std::string,int> hKw;
		std::vector<std::string> dKw;
		std::vector<int64_t> dOff { 0 };
		std::vector<int32_t> dTokKw, dTokPos;
		for ( const auto & dDoc : m_dDocs )
			for ( const auto & sText : dDoc )
			{
				std::istringstream tIn ( sText );
				std::string sTok;
				int iPos = 0;
				while ( tIn >> sTok )
				{
					std::string sLow;
					for ( char c : sTok )
						if ( isalnum ( (unsigned char)c ) )
							sLow += (char)tolower ( c );
					if ( sLow.empty() )
						continue;
					++iPos;
					auto it = hKw.find ( sLow );
					if ( it==hKw.end() )
					{
						it = hKw.emplace ( sLow, (int)dKw.size() ).first;
						dKw.push_back ( sLow );
					}
					dTokKw.push_back ( it->second );
					dTokPos.push_back ( iPos );
				}
				dOff.push_back ( (int64_t)dTokKw.size() );
			}
		std::vector<const char*> dFieldNames, dKwNames;
		for ( const auto & s : m_dFields ) dFieldNames.push_back ( s.c_str() );
		for ( const auto & s : dKw ) dKwNames.push_back ( s.c_str() );
		mgpu_build_doc_input tIn {};
		tIn.n_docs = (int)m_dDocs.size();
		tIn.n_fields = (int)m_dFields.size();
		tIn.field_names = dFieldNames.data();
		tIn.docids = m_dIds.data();
		tIn.n_keywords = (int)dKw.size();
		tIn.keywords = dKwNames.data();
		tIn.field_tok_offsets = dOff.data();
		tIn.tok_keyword = dTokKw.data();
		tIn.tok_pos = dTokPos.data();
		tIn.skiplist_block = 32;
		tIn.hit_format_inline = 1;
		char sErr[256] = "";
		if ( mgpu_build_index ( sPrefix.c_str(), &tIn, sErr, sizeof(sErr) )!=MGPU_OK )
		{
			sError = sErr;
			return false;
		}
		return true;
	}
};

int main ( int argc, char ** argv )
{
	const bool bExpectNoDevice = argc>2 && !strcmp ( argv[2], "--expect-no-device" );
	const std::string sDir = argc>1 ? argv[1] : "/tmp";
	std::string sError;

	// RTN.WeightBoundary: one document, `@title cat` -> rowid 0, weight 1500
	Corpus_t tCats;
	tCats.m_dFields = { "title", "content" };
	tCats.m_dIds = { 1 };
	tCats.m_dDocs = { { "If I were a cat...", "We are the greatest cat" } };
	CHECK ( tCats.Build ( sDir+"/cats", sError ) );

	GpuIndex_c tIndex;
	if ( !tIndex.Prealloc ( ( sDir+"/cats" ).c_str() ) )
	{
		if ( bExpectNoDevice && tIndex.GetLastStatus()==MGPU_E_NO_DEVICE )
		{
			printf ( "NO_DEVICE: %s\n", tIndex.GetLastError().c_str() );	// the product has no CPU fallback
			return 0;
		}
		fprintf ( stderr, "Prealloc failed: %s\n", tIndex.GetLastError().c_str() );
		return 1;
	}
	CHECK ( !bExpectNoDevice );
	{
		GpuQuery_t tQuery;
		tQuery.m_pRoot = GpuXQNode_t::Keyword ( "cat", 1, 1u<<0 );	// @title cat
		GpuMatchQueue_c tSorter ( 1000 );
		GpuQueryResultMeta_t tMeta;
		CHECK ( tIndex.MultiQuery ( tMeta, tQuery, &tSorter ) );
		CHECK ( tSorter.GetLength()==1 );
		CHECK ( tSorter.GetTotalCount()==1 );
		GpuMatch_t tMatch;
		tSorter.Flatten ( &tMatch );
		CHECK ( tMatch.m_tRowID==0 );
		CHECK ( tMatch.m_iWeight==1500 );
		CHECK ( tMatch.m_iDocID==1 );
		CHECK ( tMeta.m_dWordStats.size()==1 && tMeta.m_dWordStats[0].m_iDocs==1 && tMeta.m_dWordStats[0].m_iHits==2 );
	}

	{
		// the same test the way the reference writes it: the query TEXT goes in (tQuery.m_sQuery = "@title cat", sphParseExtendedQuery inside)
		GpuQuery_t tQuery;
		tQuery.m_sQuery = "@title cat";
		GpuMatchQueue_c tSorter ( 1000 );
		GpuQueryResultMeta_t tMeta;
		CHECK ( tIndex.MultiQuery ( tMeta, tQuery, &tSorter ) );
		CHECK ( tSorter.GetLength()==1 && tSorter.GetTotalCount()==1 );
		GpuMatch_t tMatch;
		tSorter.Flatten ( &tMatch );
		CHECK ( tMatch.m_tRowID==0 && tMatch.m_iWeight==1500 && tMatch.m_iDocID==1 );
		CHECK ( tMeta.m_dWordStats.size()==1 && tMeta.m_dWordStats[0].m_sWord=="cat" && tMeta.m_dWordStats[0].m_iHits==2 );

		// a text the parser refuses: MultiQuery returns false with the parser's message, nothing is pushed
		GpuQuery_t tBad;
		tBad.m_sQuery = "cat | (";
		GpuMatchQueue_c tSorter2 ( 10 );
		GpuQueryResultMeta_t tMeta2;
		CHECK ( !tIndex.MultiQuery ( tMeta2, tBad, &tSorter2 ) );
		CHECK ( tSorter2.GetLength()==0 && tMeta2.m_sError.find ( "syntax error" )!=std::string::npos );

		// legacy match mode through the text path: mode=any is "we cat"/1 under SPH_RANK_MATCHANY (PrepareQueryEmulation)
		GpuQuery_t tAny;
		tAny.m_sQuery = "dog cat";
		tAny.m_eMode = MGPU_MATCH_ANY;
		GpuMatchQueue_c tSorter3 ( 10 );
		GpuQueryResultMeta_t tMeta3;
		CHECK ( tIndex.MultiQuery ( tMeta3, tAny, &tSorter3 ) );
		CHECK ( tSorter3.GetLength()==1 );
	}

	// test/test_322 "field weights": match('program flow'), field_weights=(body=2,spam=-10) -> 3:-4574, 2:-6579, 1:-14585
	Corpus_t tW;
	tW.m_dFields = { "title", "body", "spam" };
	tW.m_dIds = { 1, 2, 3, 100 };
	tW.m_dDocs = { { "sample program", "program flow direct", "sample program flow" },
		{ "one sample program", "program rev flow", "one rev flow" },
		{ "sample two program", "sub program flow", "two sub program" },
		{ "unsigned", "", "" } };
	CHECK ( tW.Build ( sDir+"/weights", sError ) );
	GpuIndex_c tIndex2;
	CHECK ( tIndex2.Prealloc ( ( sDir+"/weights" ).c_str() ) );
	{
		GpuQuery_t dQueries[2];
		for ( int i=0; i<2; ++i )
		{
			std::vector<std::unique_ptr<GpuXQNode_t>> dKids;
			dKids.push_back ( GpuXQNode_t::Keyword ( "program", 1 ) );
			dKids.push_back ( GpuXQNode_t::Keyword ( "flow", 2 ) );
			dQueries[i].m_pRoot = GpuXQNode_t::Op ( MGPU_OP_AND, std::move ( dKids ) );
			dQueries[i].m_eRanker = MGPU_RANK_PROXIMITY_BM25;
		}
		dQueries[0].m_dFieldWeights = { 1, 2, -10 };
		dQueries[1].m_dFieldWeights = { 1, 2, 1 };
		GpuMatchQueue_c tS0 ( 1000 ), tS1 ( 1000 );
		GpuMatchSorter_i * dSorters[2] = { &tS0, &tS1 };
		GpuQueryResultMeta_t dMeta[2];
		CHECK ( tIndex2.MultiQueryEx ( 2, dQueries, dMeta, dSorters ) );
		GpuMatch_t dM[4];
		CHECK ( tS0.Flatten ( dM )==3 );
		CHECK ( dM[0].m_iDocID==3 && dM[0].m_iWeight==-4574 );
		CHECK ( dM[1].m_iDocID==2 && dM[1].m_iWeight==-6579 );
		CHECK ( dM[2].m_iDocID==1 && dM[2].m_iWeight==-14585 );
		CHECK ( tS1.Flatten ( dM )==3 );
		CHECK ( dM[0].m_iDocID==1 && dM[0].m_iWeight==7415 );
		CHECK ( dM[1].m_iDocID==3 && dM[1].m_iWeight==6426 );
		CHECK ( dM[2].m_iDocID==2 && dM[2].m_iWeight==4421 );
		CHECK ( dMeta[0].m_iTotalMatches==3 );
	}

	// error behaviour: an operator the kernels do not implement -> MultiQuery returns false with an error, no fallback
	{
		GpuQuery_t tQuery;
		std::vector<std::unique_ptr<GpuXQNode_t>> dKids;
		dKids.push_back ( GpuXQNode_t::Keyword ( "program", 1 ) );
		dKids.push_back ( GpuXQNode_t::Keyword ( "flow", 2 ) );
		dKids.push_back ( GpuXQNode_t::Keyword ( "sample", 3 ) );
		tQuery.m_pRoot = GpuXQNode_t::Op ( MGPU_OP_NEAR, std::move ( dKids ) );	// n-way NEAR: its FSM is stateful across documents in the reference
		GpuMatchQueue_c tSorter ( 10 );
		GpuQueryResultMeta_t tMeta;
		CHECK ( !tIndex2.MultiQuery ( tMeta, tQuery, &tSorter ) );
		CHECK ( !tMeta.m_sError.empty() );
		CHECK ( tSorter.GetLength()==0 );
	}
	printf ( "boundary OK\n" );
	return 0;
}
