// The C ABI of include/mgpu_writer.h: the index writer + synthetic corpus as a host-only library (libmgpu_writer.so).
#include "index_writer.h"
#include "../../../include/mgpu_writer.h"

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

using namespace mgpu;

static void CopyErr ( char * szErr, int iLen, const std::string & s )
{
	if ( szErr && iLen>0 )
		snprintf ( szErr, (size_t)iLen, "%s", s.c_str() );
}

extern "C"
{

int mgpu_writer_abi_version ( void )
{
	return 1;
}

int mgpu_build_index ( const char * path_prefix, const mgpu_build_doc_input * in, char * err, int errlen )
{
	if ( !path_prefix || !in )
		return MGPU_E_BAD_QUERY;
	std::string sError;
	if ( !BuildIndexFromDocs ( path_prefix, *in, sError ) )
	{
		CopyErr ( err, errlen, sError );
		return MGPU_E_IO;
	}
	return MGPU_OK;
}

int mgpu_build_synthetic ( const char * path_prefix, const mgpu_synth_params * p, char * err, int errlen )
{
	if ( !path_prefix || !p )
		return MGPU_E_BAD_QUERY;
	std::string sError;
	if ( !BuildSyntheticIndex ( path_prefix, *p, sError ) )
	{
		CopyErr ( err, errlen, sError );
		return MGPU_E_IO;
	}
	return MGPU_OK;
}

// the corpus object is rebuilt when the parameters change; cached per thread for the query generators
static thread_local std::unique_ptr<SynthCorpus_c> g_pCorpus;
static const SynthCorpus_c & GetCorpus ( const mgpu_synth_params * p )
{
	if ( !g_pCorpus || memcmp ( &g_pCorpus->m_tP, p, sizeof(*p) )!=0 )
		g_pCorpus.reset ( new SynthCorpus_c ( *p ) );
	return *g_pCorpus;
}

int32_t mgpu_synth_field_len ( const mgpu_synth_params * p, int64_t doc, int field )
{
	return p ? GetCorpus ( p ).FieldLen ( doc, field ) : 0;
}

int32_t mgpu_synth_token ( const mgpu_synth_params * p, int64_t doc, int field, int pos0 )
{
	return p ? GetCorpus ( p ).Token ( doc, field, pos0 ) : -1;
}

} // extern "C"
