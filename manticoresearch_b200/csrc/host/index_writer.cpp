// Index format v62 writer + seeded synthetic corpus builder. See index_writer.h for citations.
#include "index_writer.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <thread>

namespace mgpu
{

//////////////////////////////////////////////////////////////////////////
// per-keyword encoder
//////////////////////////////////////////////////////////////////////////

namespace
{

/// sink that writes hitlist varints
struct SppWrite_t
{
	ByteBuf_t & m_t;
	explicit SppWrite_t ( ByteBuf_t & t ) : m_t ( t ) {}
	int64_t	Pos() const				{ return m_t.Pos(); }
	void	Zip ( uint64_t v )		{ m_t.Zip ( v ); }
	void	SeekTo ( int64_t p )	{ m_t.Truncate ( p ); }
};

/// sink that only counts hitlist bytes (pass 1 of the parallel builder)
struct SppCount_t
{
	int64_t m_iPos = 0;
	int64_t	Pos() const				{ return m_iPos; }
	void	Zip ( uint64_t v )		{ m_iPos += ZippedLen ( v ); }
	void	SeekTo ( int64_t p )	{ m_iPos = p; }
};

struct Skip_t { DWORD m_uBaseRowIDPlus1; int64_t m_iOffset; int64_t m_iBaseHitlistPos; };

/// the state machine of CSphHitBuilder::cidxHit for a single keyword.
/// SPP: hitlist sink. If pSpd is null only the hitlist stream is produced/counted.
template<typename SPP>
TermOut_t EncodeTerm ( const RowID_t * pRows, const Hitpos_t * pHits, int64_t iCount, int iBlk, bool bInline,
	ByteBuf_t * pSpd, int64_t iSpdBase, SPP & tSpp, int64_t iSppBase, ByteBuf_t * pSpe, int64_t * pSkiplistLocal, std::vector<Skip_t> & dSkiplist )
{
	TermOut_t tRes;
	dSkiplist.clear();

	RowID_t tLastRowID = INVALID_ROWID;	// m_tLastHit.m_tRowID after HitReset()
	int64_t iLastHitlistPos = 0;		// m_iLastHitlistPos, absolute; reset per keyword (:8606)
	const int64_t iDoclistOffset = pSpd ? iSpdBase+pSpd->Pos() : 0;

	int64_t i = 0;
	while ( i<iCount )
	{
		const RowID_t tRowID = pRows[i];

		// DoclistBeginEntry, :8441-8458
		if ( pSpd )
		{
			if ( ( tRes.m_iDocs & ( iBlk-1 ) )==0 )
				dSkiplist.push_back ( { tLastRowID+1, iSpdBase+pSpd->Pos(), iLastHitlistPos } );
			pSpd->Zip ( (DWORD)( tRowID-tLastRowID ) );
		}
		const int64_t iDocHitlistPos = iSppBase+tSpp.Pos();
		const int64_t iLastHitlistDelta = iDocHitlistPos-iLastHitlistPos;
		tLastRowID = tRowID;
		iLastHitlistPos = iDocHitlistPos;

		// the hits of this document, :8629-8718
		Hitpos_t uLastWordPos = EMPTY_HIT;	// m_tLastHit.m_iWordPos
		Hitpos_t uPrevHitPos = 0;			// m_iPrevHitPos
		bool bGotFieldEnd = false;
		DWORD uDocHits = 0;
		DWORD uDocFields = 0;
		for ( ; i<iCount && pRows[i]==tRowID; ++i )
		{
			const Hitpos_t uHit = pHits[i];
			const Hitpos_t uPure = HITMAN::GetPosWithField ( uHit );
			if ( uPure==uLastWordPos )
				continue;	// duplicate position, keep the 1st

			if ( bGotFieldEnd )
			{
				if ( HITMAN::GetField ( uHit )!=HITMAN::GetField ( uLastWordPos ) )	// is the field end flag real?
					uLastWordPos |= HITMAN::FIELDEND_MASK;
				tSpp.Zip ( uLastWordPos-uPrevHitPos );
				bGotFieldEnd = false;
			}

			if ( uPure==uHit )
			{
				tSpp.Zip ( uHit-uLastWordPos );
				uLastWordPos = uHit;
			} else
			{
				bGotFieldEnd = true;
				uPrevHitPos = uLastWordPos;
				uLastWordPos = uPure;
			}

			int iField = HITMAN::GetField ( uHit );
			if ( iField<32 )
				uDocFields |= 1u<<iField;
			uDocHits++;
			tRes.m_iHits++;
		}

		// next doc / next word flush, :8567-8590
		if ( bGotFieldEnd )
		{
			uLastWordPos |= HITMAN::FIELDEND_MASK;
			tSpp.Zip ( uLastWordPos-uPrevHitPos );
		}
		const Hitpos_t uLastPos = uLastWordPos;
		if ( uLastWordPos!=EMPTY_HIT )
			tSpp.Zip ( 0 );

		// DoclistEndEntry, :8461-8497
		if ( bInline )
		{
			if ( pSpd )
				pSpd->Zip ( uDocHits );
			if ( uDocHits==1 )
			{
				tSpp.SeekTo ( iLastHitlistPos-iSppBase );	// the only hit lives in the doclist, drop it from .spp
				if ( pSpd )
				{
					pSpd->Zip ( uLastPos & 0x7FFFFF );
					pSpd->Zip ( uLastPos >> 23 );
				}
				iLastHitlistPos -= iLastHitlistDelta;
			} else if ( pSpd )
			{
				pSpd->Zip ( uDocFields );
				pSpd->Zip ( (uint64_t)iLastHitlistDelta );
			}
		} else if ( pSpd )
		{
			pSpd->Zip ( (uint64_t)iLastHitlistDelta );
			pSpd->Zip ( uDocFields );
			pSpd->Zip ( uDocHits );
		}
		tRes.m_iDocs++;
	}

	if ( !pSpd )
		return tRes;

	// DoclistEndList, :8500-8542
	pSpd->Zip ( 0 );
	if ( pSkiplistLocal )
		*pSkiplistLocal = -1;
	if ( tRes.m_iDocs>iBlk )
	{
		if ( pSkiplistLocal )
			*pSkiplistLocal = pSpe->Pos();
		Skip_t tLast = dSkiplist[0];
		(void)iDoclistOffset;
		for ( size_t k=1; k<dSkiplist.size(); ++k )
		{
			const Skip_t & t = dSkiplist[k];
			pSpe->Zip ( t.m_uBaseRowIDPlus1-tLast.m_uBaseRowIDPlus1-(DWORD)iBlk );
			pSpe->Zip ( (uint64_t)( t.m_iOffset-tLast.m_iOffset-4*iBlk ) );
			pSpe->Zip ( (uint64_t)( t.m_iBaseHitlistPos-tLast.m_iBaseHitlistPos ) );
			tLast = t;
		}
	}
	return tRes;
}

} // namespace


TermOut_t TermEncoder_c::Encode ( const RowID_t * pRows, const Hitpos_t * pHits, int64_t iCount,
	ByteBuf_t & tSpd, int64_t iSpdBase, ByteBuf_t & tSpp, int64_t iSppBase, ByteBuf_t & tSpe, int64_t * pSkiplistLocal )
{
	std::vector<Skip_t> dSkips;
	SppWrite_t tSink ( tSpp );
	return EncodeTerm ( pRows, pHits, iCount, m_iBlk, m_bInline, &tSpd, iSpdBase, tSink, iSppBase, &tSpe, pSkiplistLocal, dSkips );
}

//////////////////////////////////////////////////////////////////////////
// files
//////////////////////////////////////////////////////////////////////////

bool WriteFile ( const std::string & sPath, const void * pData, size_t iLen, std::string & sError )
{
	FILE * fp = fopen ( sPath.c_str(), "wb" );
	if ( !fp )
	{
		sError = "failed to open " + sPath + " for writing";
		return false;
	}
	const size_t CHUNK = (size_t)1<<30;
	const BYTE * p = (const BYTE *)pData;
	size_t iLeft = iLen;
	while ( iLeft )
	{
		size_t n = std::min ( iLeft, CHUNK );
		if ( fwrite ( p, 1, n, fp )!=n )
		{
			fclose ( fp );
			sError = "write error on " + sPath;
			return false;
		}
		p += n;
		iLeft -= n;
	}
	fclose ( fp );
	return true;
}


/// appends pieces to a file one after another (big synthetic indexes are written range by range)
struct FileAppender_c
{
	FILE * m_fp = nullptr;
	int64_t m_iPos = 0;
	bool Open ( const std::string & sPath, std::string & sError )
	{
		m_fp = fopen ( sPath.c_str(), "wb" );
		if ( !m_fp )
			sError = "failed to open " + sPath + " for writing";
		return m_fp!=nullptr;
	}
	bool Put ( const void * p, size_t n )
	{
		m_iPos += (int64_t)n;
		return !n || fwrite ( p, 1, n, m_fp )==n;
	}
	void Close()	{ if ( m_fp ) fclose ( m_fp ); m_fp = nullptr; }
	~FileAppender_c() { Close(); }
};


bool WriteAttrsAndHeader ( const std::string & sPrefix, IndexHeader_t & tHdr, const std::vector<DWORD> & dRows, int iStride, int64_t iRows, std::string & sError )
{
	// .spa = rows, then per-128-row {min row, max row}, then index-wide {min,max}: AttrIndexBuilder_c, src/sphinx.cpp:662-800
	std::vector<DWORD> dSpa ( dRows );
	const int nAttrs = (int)tHdr.m_dAttrs.size();
	auto fnGet = [&] ( const DWORD * pRow, const SchemaAttr_t & a ) -> int64_t
	{
		if ( a.m_iBitCount==64 )
			return (int64_t)( (uint64_t)pRow[a.m_iBitOffset/32] | ( (uint64_t)pRow[a.m_iBitOffset/32+1]<<32 ) );
		return (int64_t)pRow[a.m_iBitOffset/32];
	};
	auto fnSet = [&] ( DWORD * pRow, const SchemaAttr_t & a, int64_t v )
	{
		pRow[a.m_iBitOffset/32] = (DWORD)(uint64_t)v;
		if ( a.m_iBitCount==64 )
			pRow[a.m_iBitOffset/32+1] = (DWORD)( (uint64_t)v>>32 );
	};

	std::vector<int64_t> dIdxMin ( nAttrs, INT64_MAX ), dIdxMax ( nAttrs, INT64_MIN );
	int64_t nBlocks = 0;
	for ( int64_t iStart=0; iStart<iRows; iStart+=DOCINFO_INDEX_FREQ )
	{
		int64_t iEnd = std::min ( iStart+DOCINFO_INDEX_FREQ, iRows );
		std::vector<int64_t> dMin ( nAttrs, INT64_MAX ), dMax ( nAttrs, INT64_MIN );
		for ( int64_t r=iStart; r<iEnd; ++r )
			for ( int a=0; a<nAttrs; ++a )
			{
				int64_t v = fnGet ( &dRows[r*iStride], tHdr.m_dAttrs[a] );
				dMin[a] = std::min ( dMin[a], v );
				dMax[a] = std::max ( dMax[a], v );
			}
		std::vector<DWORD> dMinRow ( iStride, 0 ), dMaxRow ( iStride, 0 );
		for ( int a=0; a<nAttrs; ++a )
		{
			dIdxMin[a] = std::min ( dIdxMin[a], dMin[a] );
			dIdxMax[a] = std::max ( dIdxMax[a], dMax[a] );
			fnSet ( dMinRow.data(), tHdr.m_dAttrs[a], dMin[a] );
			fnSet ( dMaxRow.data(), tHdr.m_dAttrs[a], dMax[a] );
		}
		dSpa.insert ( dSpa.end(), dMinRow.begin(), dMinRow.end() );
		dSpa.insert ( dSpa.end(), dMaxRow.begin(), dMaxRow.end() );
		nBlocks++;
	}
	{
		std::vector<DWORD> dMinRow ( iStride, 0 ), dMaxRow ( iStride, 0 );
		for ( int a=0; a<nAttrs; ++a )
		{
			fnSet ( dMinRow.data(), tHdr.m_dAttrs[a], dIdxMin[a] );
			fnSet ( dMaxRow.data(), tHdr.m_dAttrs[a], dIdxMax[a] );
		}
		dSpa.insert ( dSpa.end(), dMinRow.begin(), dMinRow.end() );
		dSpa.insert ( dSpa.end(), dMaxRow.begin(), dMaxRow.end() );
	}

	tHdr.m_iDocinfo = iRows;
	tHdr.m_iDocinfoIndex = nBlocks;
	tHdr.m_iMinMaxIndex = iRows*iStride;

	if ( !WriteFile ( sPrefix+".spa", dSpa.data(), dSpa.size()*sizeof(DWORD), sError ) )
		return false;

	// dead-row map: one bit per row, all alive (src/killlist.*)
	std::vector<DWORD> dSpm ( (size_t)( ( iRows+31 )/32 ), 0 );
	if ( !WriteFile ( sPrefix+".spm", dSpm.data(), dSpm.size()*sizeof(DWORD), sError ) )
		return false;

	ByteBuf_t tSph;
	WriteHeader ( tSph, tHdr );
	return WriteFile ( sPrefix+".sph", tSph.m_d.data(), tSph.m_d.size(), sError );
}


static void SetupSchema ( IndexHeader_t & tHdr, int nFields, const char * const * ppFields, int nAttrs, const char * const * ppAttrs )
{
	for ( int i=0; i<nFields; ++i )
	{
		SchemaField_t f;
		f.m_sName = ppFields[i];
		tHdr.m_dFields.push_back ( f );
	}
	// the first attribute is always `id` (bigint, 2 DWORDs), src/attribute.h:142-146
	SchemaAttr_t tId;
	tId.m_sName = "id";
	tId.m_eType = SPH_ATTR_BIGINT;
	tId.m_iBitOffset = 0;
	tId.m_iBitCount = 64;
	tHdr.m_dAttrs.push_back ( tId );
	for ( int i=0; i<nAttrs; ++i )
	{
		SchemaAttr_t a;
		a.m_sName = ppAttrs[i];
		a.m_eType = SPH_ATTR_INTEGER;
		a.m_iBitOffset = 64+32*i;
		a.m_iBitCount = 32;
		tHdr.m_dAttrs.push_back ( a );
	}
}

//////////////////////////////////////////////////////////////////////////
// builder from explicit documents (tests, golden corpora)
//////////////////////////////////////////////////////////////////////////

bool BuildIndexFromDocs ( const char * szPrefix, const mgpu_build_doc_input & tIn, std::string & sError )
{
	const std::string sPrefix ( szPrefix );
	const int iBlk = tIn.skiplist_block>0 ? tIn.skiplist_block : 32;
	if ( iBlk & ( iBlk-1 ) )
	{
		sError = "skiplist block size must be a power of two";
		return false;
	}
	if ( tIn.n_fields<1 || tIn.n_fields>32 )
	{
		sError = "1..32 fields supported";
		return false;
	}

	IndexHeader_t tHdr;
	SetupSchema ( tHdr, tIn.n_fields, tIn.field_names, tIn.n_attrs, tIn.attr_names );
	tHdr.m_iSkiplistBlockSize = iBlk;
	tHdr.m_eHitFormat = tIn.hit_format_inline ? SPH_HIT_FORMAT_INLINE : SPH_HIT_FORMAT_PLAIN;
	tHdr.m_iTotalDocuments = (DWORD)tIn.n_docs;
	const bool bCrc = tIn.dict_crc!=0;
	tHdr.m_bWordDict = bCrc ? 0 : 1;

	// keyword order = strcmp order (CSphDictKeywords sorts its chunks with strcmp, src/sphinx.cpp:19589-19595);
	// dict=crc: ascending word id (the hit stream is sorted by word id, CSphHitBuilder::cidxHit sees them in that order)
	std::vector<int> dKwOrder ( tIn.n_keywords );
	std::iota ( dKwOrder.begin(), dKwOrder.end(), 0 );
	std::vector<uint64_t> dKwID ( tIn.n_keywords, 0 );
	if ( bCrc )
	{
		for ( int i=0; i<tIn.n_keywords; ++i )
			dKwID[i] = WordIdFNV64 ( tIn.keywords[i] );
		std::sort ( dKwOrder.begin(), dKwOrder.end(), [&] ( int a, int b ) { return dKwID[a]<dKwID[b]; } );
		for ( int i=0; i+1<tIn.n_keywords; ++i )
			if ( dKwID[dKwOrder[i]]==dKwID[dKwOrder[i+1]] || !dKwID[dKwOrder[i]] )
			{
				sError = "keywords collide under FNV64 (or hash to zero)";
				return false;
			}
	} else
		std::sort ( dKwOrder.begin(), dKwOrder.end(), [&] ( int a, int b ) { return strcmp ( tIn.keywords[a], tIn.keywords[b] )<0; } );
	std::vector<int> dKwRank ( tIn.n_keywords );
	for ( int i=0; i<tIn.n_keywords; ++i )
		dKwRank[dKwOrder[i]] = i;

	// the hit stream: CSphSource_Document::BuildRegularHits (src/sphinx.cpp:22437-22549) sets the field-end marker
	// on every hit sitting at the last position of a field
	struct Hit_t { int m_iKw; RowID_t m_tRow; Hitpos_t m_uPos; };
	std::vector<Hit_t> dHits;
	int64_t iTotalBytes = 0;
	for ( int iDoc=0; iDoc<tIn.n_docs; ++iDoc )
		for ( int iField=0; iField<tIn.n_fields; ++iField )
		{
			int64_t iFrom = tIn.field_tok_offsets[(int64_t)iDoc*tIn.n_fields+iField];
			int64_t iTo = tIn.field_tok_offsets[(int64_t)iDoc*tIn.n_fields+iField+1];
			int iMaxPos = 0;
			for ( int64_t k=iFrom; k<iTo; ++k )
				iMaxPos = std::max ( iMaxPos, tIn.tok_pos[k] );
			for ( int64_t k=iFrom; k<iTo; ++k )
			{
				if ( tIn.tok_keyword[k]<0 || tIn.tok_keyword[k]>=tIn.n_keywords || tIn.tok_pos[k]<1 || tIn.tok_pos[k]>(int)HITMAN::POS_MASK )
				{
					sError = "bad token";
					return false;
				}
				Hitpos_t uPos = HITMAN::Create ( iField, tIn.tok_pos[k] );
				if ( tIn.tok_pos[k]==iMaxPos )
					uPos |= HITMAN::FIELDEND_MASK;
				dHits.push_back ( { dKwRank[tIn.tok_keyword[k]], (RowID_t)iDoc, uPos } );
				iTotalBytes += (int64_t)strlen ( tIn.keywords[tIn.tok_keyword[k]] )+1;
			}
		}
	tHdr.m_iTotalBytes = iTotalBytes;
	std::sort ( dHits.begin(), dHits.end(), [] ( const Hit_t & a, const Hit_t & b )
	{
		if ( a.m_iKw!=b.m_iKw ) return a.m_iKw<b.m_iKw;
		if ( a.m_tRow!=b.m_tRow ) return a.m_tRow<b.m_tRow;
		return a.m_uPos<b.m_uPos;
	});

	ByteBuf_t tSpd, tSpp, tSpe;
	tSpd.PutByte ( 1 ); tSpp.PutByte ( 1 ); tSpe.PutByte ( 1 );	// CreateIndexFiles, :8404-8409
	DictWriter_c tDict ( iBlk, bCrc );
	TermEncoder_c tEnc ( iBlk, tIn.hit_format_inline!=0 );

	std::vector<RowID_t> dRows;
	std::vector<Hitpos_t> dPos;
	size_t i = 0;
	while ( i<dHits.size() )
	{
		size_t j = i;
		dRows.clear(); dPos.clear();
		while ( j<dHits.size() && dHits[j].m_iKw==dHits[i].m_iKw )
		{
			dRows.push_back ( dHits[j].m_tRow );
			dPos.push_back ( dHits[j].m_uPos );
			++j;
		}
		DictEntry_t tEntry;
		tEntry.m_sKeyword = tIn.keywords[dKwOrder[dHits[i].m_iKw]];
		tEntry.m_uWordID = dKwID[dKwOrder[dHits[i].m_iKw]];
		tEntry.m_iDoclistOffset = tSpd.Pos();
		int64_t iSkipLocal = -1;
		TermOut_t tOut = tEnc.Encode ( dRows.data(), dPos.data(), (int64_t)dRows.size(), tSpd, 0, tSpp, 0, tSpe, &iSkipLocal );
		tEntry.m_iDocs = tOut.m_iDocs;
		tEntry.m_iHits = tOut.m_iHits;
		tEntry.m_iDoclistLength = tSpd.Pos()-tEntry.m_iDoclistOffset;
		tEntry.m_iSkiplistOffset = iSkipLocal>=0 ? iSkipLocal : 0;
		tDict.AddEntry ( tEntry );
		i = j;
	}
	tDict.Finish ( tHdr, tSpd.Pos() );

	if ( !WriteFile ( sPrefix+".spd", tSpd.m_d.data(), tSpd.m_d.size(), sError ) ) return false;
	if ( !WriteFile ( sPrefix+".spp", tSpp.m_d.data(), tSpp.m_d.size(), sError ) ) return false;
	if ( !WriteFile ( sPrefix+".spe", tSpe.m_d.data(), tSpe.m_d.size(), sError ) ) return false;
	if ( !WriteFile ( sPrefix+".spi", tDict.m_tOut.m_d.data(), tDict.m_tOut.m_d.size(), sError ) ) return false;

	const int iStride = 2+tIn.n_attrs;
	std::vector<DWORD> dRowsAttr ( (size_t)tIn.n_docs*iStride );
	for ( int iDoc=0; iDoc<tIn.n_docs; ++iDoc )
	{
		uint64_t uId = (uint64_t)tIn.docids[iDoc];
		dRowsAttr[(size_t)iDoc*iStride] = (DWORD)uId;
		dRowsAttr[(size_t)iDoc*iStride+1] = (DWORD)( uId>>32 );
		for ( int a=0; a<tIn.n_attrs; ++a )
			dRowsAttr[(size_t)iDoc*iStride+2+a] = tIn.attrs[(size_t)iDoc*tIn.n_attrs+a];
	}
	return WriteAttrsAndHeader ( sPrefix, tHdr, dRowsAttr, iStride, tIn.n_docs, sError );
}

//////////////////////////////////////////////////////////////////////////
// seeded synthetic corpus
//////////////////////////////////////////////////////////////////////////

static inline uint64_t Mix64 ( uint64_t z )
{
	// splitmix64 finalizer
	z += 0x9E3779B97F4A7C15ull;
	z = ( z ^ ( z>>30 ) ) * 0xBF58476D1CE4E5B9ull;
	z = ( z ^ ( z>>27 ) ) * 0x94D049BB133111EBull;
	return z ^ ( z>>31 );
}

static double NormInv ( double p )
{
	// Acklam's rational approximation of the inverse normal CDF
	static const double a[] = { -3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02, 1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00 };
	static const double b[] = { -5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02, 6.680131188771972e+01, -1.328068155288572e+01 };
	static const double c[] = { -7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00, -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00 };
	static const double d[] = { 7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00 };
	const double plow = 0.02425, phigh = 1-plow;
	if ( p<plow )
	{
		double q = sqrt ( -2*log ( p ) );
		return ( ( ( ( ( c[0]*q+c[1] )*q+c[2] )*q+c[3] )*q+c[4] )*q+c[5] ) / ( ( ( ( d[0]*q+d[1] )*q+d[2] )*q+d[3] )*q+1 );
	}
	if ( p>phigh )
	{
		double q = sqrt ( -2*log ( 1-p ) );
		return -( ( ( ( ( c[0]*q+c[1] )*q+c[2] )*q+c[3] )*q+c[4] )*q+c[5] ) / ( ( ( ( d[0]*q+d[1] )*q+d[2] )*q+d[3] )*q+1 );
	}
	double q = p-0.5, r = q*q;
	return ( ( ( ( ( a[0]*r+a[1] )*r+a[2] )*r+a[3] )*r+a[4] )*r+a[5] )*q / ( ( ( ( ( b[0]*r+b[1] )*r+b[2] )*r+b[3] )*r+b[4] )*r+1 );
}

SynthCorpus_c::SynthCorpus_c ( const mgpu_synth_params & p )
	: m_tP ( p )
{
	const int V = p.vocab;
	// Zipf(s=1) over ranks 1..V, Vose's alias method
	std::vector<double> dP ( V );
	double fH = 0;
	for ( int r=1; r<=V; ++r )
		fH += 1.0/r;
	for ( int r=1; r<=V; ++r )
		dP[r-1] = ( 1.0/r )/fH*V;
	m_dAliasProb.assign ( V, 0xFFFFFFFFu );
	m_dAlias.resize ( V );
	std::vector<int> dSmall, dLarge;
	for ( int i=0; i<V; ++i )
	{
		m_dAlias[i] = i;
		( dP[i]<1.0 ? dSmall : dLarge ).push_back ( i );
	}
	while ( !dSmall.empty() && !dLarge.empty() )
	{
		int s = dSmall.back(); dSmall.pop_back();
		int l = dLarge.back(); dLarge.pop_back();
		double f = dP[s]*4294967296.0;
		m_dAliasProb[s] = f>=4294967295.0 ? 0xFFFFFFFFu : (uint32_t)f;
		m_dAlias[s] = l;
		dP[l] = ( dP[l]+dP[s] )-1.0;
		( dP[l]<1.0 ? dSmall : dLarge ).push_back ( l );
	}

	m_dBodyLenTable.resize ( 4096 );
	for ( int i=0; i<4096; ++i )
	{
		double z = NormInv ( ( i+0.5 )/4096.0 );
		double f = exp ( (double)p.body_mu + (double)p.body_sigma*z );
		int n = (int)floor ( f+0.5 );
		n = std::max ( p.body_min, std::min ( p.body_max, n ) );
		m_dBodyLenTable[i] = (uint16_t)n;
	}
}

int SynthCorpus_c::FieldLen ( int64_t iDoc, int iField ) const
{
	uint64_t h = Mix64 ( m_tP.seed ^ Mix64 ( (uint64_t)iDoc*2+(uint64_t)iField+0x51ull ) );
	if ( iField==0 )
		return m_tP.title_min + (int)( ( h>>33 ) % (uint64_t)( m_tP.title_max-m_tP.title_min+1 ) );
	return m_dBodyLenTable [ ( h>>40 ) & 4095 ];
}

int SynthCorpus_c::Token ( int64_t iDoc, int iField, int iPos0 ) const
{
	uint64_t h = Mix64 ( ( m_tP.seed+0x1234567ull ) ^ Mix64 ( ( (uint64_t)iDoc<<12 ) ^ ( (uint64_t)iField<<11 ) ^ (uint64_t)iPos0 ) );
	uint32_t uIdx = (uint32_t)( ( ( h>>32 ) * (uint64_t)m_tP.vocab )>>32 );
	uint32_t uFrac = (uint32_t)h;
	return uFrac<m_dAliasProb[uIdx] ? (int)uIdx : (int)m_dAlias[uIdx];
}

DWORD SynthCorpus_c::AttrGid ( int64_t iDoc ) const
{
	return (DWORD)( Mix64 ( (uint64_t)( iDoc+1 ) ) % 1000 );
}

DWORD SynthCorpus_c::AttrTs ( int64_t iDoc ) const
{
	return 1500000000u + (DWORD)( Mix64 ( (uint64_t)( iDoc+1 ) ^ 1 ) % 100000000ull );
}


template<typename FN>
static void ParallelFor ( int nThreads, FN && fn )
{
	std::vector<std::thread> dThreads;
	for ( int t=1; t<nThreads; ++t )
		dThreads.emplace_back ( [&fn,t] { fn ( t ); } );
	fn ( 0 );
	for ( auto & t : dThreads )
		t.join();
}


bool BuildSyntheticIndex ( const char * szPrefix, const mgpu_synth_params & tParams, std::string & sError )
{
	const std::string sPrefix ( szPrefix );
	const int iBlk = 32;
	const int64_t nDocs = tParams.n_docs;
	const int V = tParams.vocab;
	if ( nDocs<1 || nDocs>=(int64_t)INVALID_ROWID || V<2 || V>( 1<<24 ) || tParams.body_max>4000 || tParams.title_max>2000 )
	{
		sError = "bad synthetic corpus parameters";
		return false;
	}
	int nThreads = tParams.threads>0 ? tParams.threads : (int)std::thread::hardware_concurrency();
	nThreads = std::max ( 1, std::min ( nThreads, 256 ) );
	if ( (int64_t)nThreads>nDocs )
		nThreads = (int)nDocs;

	SynthCorpus_c tCorpus ( tParams );

	// ---- phase A: per-thread per-term hit counts over contiguous doc ranges
	std::vector<std::vector<uint32_t>> dCounts ( nThreads );
	auto fnDocRange = [&] ( int t, int64_t & iFrom, int64_t & iTo ) { iFrom = nDocs*t/nThreads; iTo = nDocs*( t+1 )/nThreads; };
	std::vector<int64_t> dBytesPerThread ( nThreads, 0 );
	ParallelFor ( nThreads, [&] ( int t )
	{
		auto & dC = dCounts[t];
		dC.assign ( V, 0 );
		int64_t iFrom, iTo;
		fnDocRange ( t, iFrom, iTo );
		int64_t iBytes = 0;
		for ( int64_t d=iFrom; d<iTo; ++d )
		{
			int64_t iDoc = tParams.first_doc+d;
			for ( int f=0; f<2; ++f )
			{
				int n = tCorpus.FieldLen ( iDoc, f );
				for ( int k=0; k<n; ++k )
					dC [ tCorpus.Token ( iDoc, f, k ) ]++;
				iBytes += 9*(int64_t)n;
			}
		}
		dBytesPerThread[t] = iBytes;
	});

	// ---- term start offsets (term-major), and per-thread write cursors
	std::vector<int64_t> dTermStart ( (size_t)V+1 );
	{
		int64_t iAcc = 0;
		for ( int w=0; w<V; ++w )
		{
			dTermStart[w] = iAcc;
			for ( int t=0; t<nThreads; ++t )
			{
				uint32_t c = dCounts[t][w];
				// reuse the count slot as "offset of this thread inside the term"; 32 bits suffice per term
				dCounts[t][w] = (uint32_t)( iAcc-dTermStart[w] );
				iAcc += c;
			}
		}
		dTermStart[V] = iAcc;
	}
	const int64_t nHits = dTermStart[V];
	for ( int w=0; w<V; ++w )
		if ( dTermStart[w+1]-dTermStart[w]>=( (int64_t)1<<32 ) )
		{
			sError = "term too frequent for 32-bit per-term cursors";
			return false;
		}

	// ---- phase B: stable scatter (docs ascending inside each thread, threads own ascending doc ranges)
	std::vector<RowID_t> dRows ( (size_t)nHits );
	std::vector<Hitpos_t> dPos ( (size_t)nHits );
	ParallelFor ( nThreads, [&] ( int t )
	{
		auto & dC = dCounts[t];
		int64_t iFrom, iTo;
		fnDocRange ( t, iFrom, iTo );
		for ( int64_t d=iFrom; d<iTo; ++d )
		{
			int64_t iDoc = tParams.first_doc+d;
			for ( int f=0; f<2; ++f )
			{
				int n = tCorpus.FieldLen ( iDoc, f );
				for ( int k=0; k<n; ++k )
				{
					int w = tCorpus.Token ( iDoc, f, k );
					int64_t iAt = dTermStart[w] + dC[w]++;
					dRows[iAt] = (RowID_t)d;
					Hitpos_t uPos = HITMAN::Create ( f, k+1 );
					if ( k==n-1 )
						uPos |= HITMAN::FIELDEND_MASK;
					dPos[iAt] = uPos;
				}
			}
		}
	});
	dCounts.clear();
	dCounts.shrink_to_fit();

	// ---- phase C: encode. Terms are cut into contiguous ranges of roughly equal hit counts.
	const int nRanges = nThreads*8;
	std::vector<int> dRangeStart ( nRanges+1 );
	{
		int w = 0;
		for ( int r=0; r<nRanges; ++r )
		{
			dRangeStart[r] = w;
			int64_t iTarget = nHits*( r+1 )/nRanges;
			while ( w<V && dTermStart[w+1]<=iTarget )
				++w;
			if ( r==nRanges-1 )
				w = V;
		}
		dRangeStart[nRanges] = V;
		for ( int r=1; r<=nRanges; ++r )
			dRangeStart[r] = std::max ( dRangeStart[r], dRangeStart[r-1] );
	}

	// pass 1: hitlist bytes per range (hitlist content does not depend on absolute offsets)
	std::vector<int64_t> dRangeSpp ( nRanges+1, 0 );
	std::atomic<int> iNext { 0 };
	ParallelFor ( nThreads, [&] ( int )
	{
		std::vector<Skip_t> dSkips;
		for ( int r; ( r = iNext.fetch_add(1) )<nRanges; )
		{
			SppCount_t tCnt;
			for ( int w=dRangeStart[r]; w<dRangeStart[r+1]; ++w )
			{
				int64_t n = dTermStart[w+1]-dTermStart[w];
				if ( n )
					EncodeTerm ( &dRows[dTermStart[w]], &dPos[dTermStart[w]], n, iBlk, true, nullptr, 0, tCnt, 0, nullptr, nullptr, dSkips );
			}
			dRangeSpp[r] = tCnt.m_iPos;
		}
	});
	std::vector<int64_t> dSppBase ( nRanges+1 );
	dSppBase[0] = 1;	// dummy byte
	for ( int r=0; r<nRanges; ++r )
		dSppBase[r+1] = dSppBase[r]+dRangeSpp[r];

	// pass 2: full encode per range, written out in range order
	struct RangeOut_t
	{
		ByteBuf_t m_tSpd, m_tSpp, m_tSpe;
		std::vector<DictEntry_t> m_dEntries;	// offsets local to the range buffers
		bool m_bDone = false;
	};
	std::vector<RangeOut_t> dOut ( nRanges );
	std::vector<std::atomic<int>> dReady ( nRanges );
	for ( auto & a : dReady )
		a.store ( 0 );

	FileAppender_c tFSpd, tFSpp, tFSpe;
	if ( !tFSpd.Open ( sPrefix+".spd", sError ) || !tFSpp.Open ( sPrefix+".spp", sError ) || !tFSpe.Open ( sPrefix+".spe", sError ) )
		return false;
	BYTE bDummy = 1;
	tFSpd.Put ( &bDummy, 1 ); tFSpp.Put ( &bDummy, 1 ); tFSpe.Put ( &bDummy, 1 );

	DictWriter_c tDict ( iBlk );
	IndexHeader_t tHdr;
	const char * dFieldNames[] = { "title", "body" };
	const char * dAttrNames[] = { "gid", "ts" };
	SetupSchema ( tHdr, 2, dFieldNames, 2, dAttrNames );
	tHdr.m_iSkiplistBlockSize = iBlk;
	tHdr.m_eHitFormat = SPH_HIT_FORMAT_INLINE;
	tHdr.m_iTotalDocuments = (DWORD)nDocs;
	tHdr.m_iTotalBytes = std::accumulate ( dBytesPerThread.begin(), dBytesPerThread.end(), (int64_t)0 );

	iNext.store ( 0 );
	bool bWriteOk = true;
	std::atomic<int> iWritten { 0 };
	// thread 0 doubles as the in-order writer so that memory for finished ranges is released early
	auto fnWriter = [&] ()
	{
		int r = iWritten.load();
		while ( r<nRanges && dReady[r].load ( std::memory_order_acquire ) )
		{
			RangeOut_t & o = dOut[r];
			const int64_t iSpdBase = tFSpd.m_iPos, iSpeBase = tFSpe.m_iPos;
			for ( auto & e : o.m_dEntries )
			{
				e.m_iDoclistOffset += iSpdBase;
				if ( e.m_iDocs>iBlk )
					e.m_iSkiplistOffset += iSpeBase;
				tDict.AddEntry ( e );
			}
			bWriteOk &= tFSpd.Put ( o.m_tSpd.m_d.data(), o.m_tSpd.m_d.size() );
			bWriteOk &= tFSpp.Put ( o.m_tSpp.m_d.data(), o.m_tSpp.m_d.size() );
			bWriteOk &= tFSpe.Put ( o.m_tSpe.m_d.data(), o.m_tSpe.m_d.size() );
			o = RangeOut_t();
			iWritten.store ( ++r );
		}
	};
	ParallelFor ( nThreads, [&] ( int t )
	{
		std::vector<Skip_t> dSkips;
		char sKw[16];
		for ( int r; ( r = iNext.fetch_add(1) )<nRanges; )
		{
			RangeOut_t & o = dOut[r];
			SppWrite_t tSink ( o.m_tSpp );
			for ( int w=dRangeStart[r]; w<dRangeStart[r+1]; ++w )
			{
				int64_t n = dTermStart[w+1]-dTermStart[w];
				if ( !n )
					continue;
				DictEntry_t e;
				snprintf ( sKw, sizeof(sKw), "t%07d", w+1 );
				e.m_sKeyword = sKw;
				e.m_iDoclistOffset = o.m_tSpd.Pos();
				int64_t iSkipLocal = -1;
				// iSpdBase=0: skiplist deltas and doclist record contents are position independent;
				// absolute doclist/skiplist offsets are fixed up by the writer
				TermOut_t tOut = EncodeTerm ( &dRows[dTermStart[w]], &dPos[dTermStart[w]], n, iBlk, true, &o.m_tSpd, 0, tSink, dSppBase[r], &o.m_tSpe, &iSkipLocal, dSkips );
				e.m_iDocs = tOut.m_iDocs;
				e.m_iHits = tOut.m_iHits;
				e.m_iDoclistLength = o.m_tSpd.Pos()-e.m_iDoclistOffset;
				e.m_iSkiplistOffset = iSkipLocal>=0 ? iSkipLocal : 0;
				o.m_dEntries.push_back ( e );
			}
			dReady[r].store ( 1, std::memory_order_release );
			if ( t==0 )
				fnWriter();
		}
	});
	fnWriter();
	if ( iWritten.load()!=nRanges || !bWriteOk )
	{
		sError = "failed to write index data files";
		return false;
	}
	if ( tFSpp.m_iPos!=dSppBase[nRanges] )
	{
		sError = "internal error: hitlist size mismatch between passes";
		return false;
	}
	tFSpd.Close(); tFSpp.Close(); tFSpe.Close();

	tDict.Finish ( tHdr );
	if ( !WriteFile ( sPrefix+".spi", tDict.m_tOut.m_d.data(), tDict.m_tOut.m_d.size(), sError ) )
		return false;

	dRows.clear(); dRows.shrink_to_fit();
	dPos.clear(); dPos.shrink_to_fit();

	const int iStride = 4;
	std::vector<DWORD> dAttrRows ( (size_t)nDocs*iStride );
	ParallelFor ( nThreads, [&] ( int t )
	{
		int64_t iFrom, iTo;
		fnDocRange ( t, iFrom, iTo );
		for ( int64_t d=iFrom; d<iTo; ++d )
		{
			int64_t iDoc = tParams.first_doc+d;
			uint64_t uId = (uint64_t)( iDoc+1 );
			dAttrRows[(size_t)d*iStride] = (DWORD)uId;
			dAttrRows[(size_t)d*iStride+1] = (DWORD)( uId>>32 );
			dAttrRows[(size_t)d*iStride+2] = tCorpus.AttrGid ( iDoc );
			dAttrRows[(size_t)d*iStride+3] = tCorpus.AttrTs ( iDoc );
		}
	});
	return WriteAttrsAndHeader ( sPrefix, tHdr, dAttrRows, iStride, nDocs, sError );
}

} // namespace mgpu
