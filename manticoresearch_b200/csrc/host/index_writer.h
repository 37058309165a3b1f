// Index format v62 writer: doclists (.spd), hitlists (.spp), skiplists (.spe), dictionary (.spi),
// attributes (.spa), header (.sph), dead-row map (.spm).
// Byte layout follows CSphHitBuilder (src/sphinx.cpp:8297-8719): DoclistBeginEntry :8441-8458,
// DoclistEndEntry :8461-8497, DoclistEndList :8500-8542, cidxHit :8554-8719.
#pragma once

#include "index_format.h"
#include "../../../include/mgpu.h"
#include "../../../include/mgpu_writer.h"

namespace mgpu
{

/// one posting list being encoded; output of TermEncoder_c::Encode
struct TermOut_t
{
	int		m_iDocs = 0;
	int		m_iHits = 0;
};

/// encodes ONE keyword's hits (sorted by rowid, then raw hitpos incl. end-marker bit) into the three
/// streams. Hitlist offsets written into the doclist are absolute .spp offsets, so the caller passes
/// the absolute position of tSpp's first byte in iSppBase (ditto iSpdBase for skiplist/dict offsets).
class TermEncoder_c
{
public:
	TermEncoder_c ( int iSkiplistBlockSize, bool bInlineHits )
		: m_iBlk ( iSkiplistBlockSize ), m_bInline ( bInlineHits ) {}

	/// pRows/pHits: iCount hits of this keyword. Appends to tSpd/tSpp/tSpe.
	/// returns docs/hits; *pSkiplistLocal = offset inside tSpe where this term's skiplist starts (or -1)
	TermOut_t	Encode ( const RowID_t * pRows, const Hitpos_t * pHits, int64_t iCount,
					ByteBuf_t & tSpd, int64_t iSpdBase, ByteBuf_t & tSpp, int64_t iSppBase, ByteBuf_t & tSpe, int64_t * pSkiplistLocal );

private:
	int		m_iBlk;
	bool	m_bInline;
};

bool	WriteFile ( const std::string & sPath, const void * pData, size_t iLen, std::string & sError );

/// writes .spa (+min-max rows), .spm (all alive), .sph; pAttrs rows are [id lo, id hi, attr0, attr1...]
bool	WriteAttrsAndHeader ( const std::string & sPrefix, IndexHeader_t & tHdr, const std::vector<DWORD> & dRows, int iStride, int64_t iRows, std::string & sError );

bool	BuildIndexFromDocs ( const char * szPrefix, const mgpu_build_doc_input & tIn, std::string & sError );
bool	BuildSyntheticIndex ( const char * szPrefix, const mgpu_synth_params & tParams, std::string & sError );

/// the seeded synthetic corpus; integer-only sampling so every builder agrees
struct SynthCorpus_c
{
	explicit SynthCorpus_c ( const mgpu_synth_params & p );
	int		FieldLen ( int64_t iDoc, int iField ) const;
	int		Token ( int64_t iDoc, int iField, int iPos0 ) const;	///< 0-based term rank-1 (0 = most frequent)
	DWORD	AttrGid ( int64_t iDoc ) const;
	DWORD	AttrTs ( int64_t iDoc ) const;

	mgpu_synth_params		m_tP;
	std::vector<uint32_t>	m_dAliasProb;	///< alias method: threshold (32-bit fixed point)
	std::vector<uint32_t>	m_dAlias;
	std::vector<uint16_t>	m_dBodyLenTable;	///< 4096-entry quantile table of the clipped lognormal
};

} // namespace mgpu
