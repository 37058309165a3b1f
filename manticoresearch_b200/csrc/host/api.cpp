// The C ABI of include/mgpu.h. Thin: argument checks, handle lifetime, error strings; no logic.
#include "engine.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>

using namespace mgpu;

struct mgpu_index { Index_c m_t; };
struct mgpu_batch { Batch_c m_t; };

static thread_local std::string g_sLastOpenError;

static void CopyErr ( char * szErr, int iLen, const std::string & s )
{
	if ( szErr && iLen>0 )
		snprintf ( szErr, (size_t)iLen, "%s", s.c_str() );
}

extern "C"
{

int mgpu_abi_version ( void )
{
	return MGPU_ABI_VERSION;
}

int mgpu_index_open ( const char * path_prefix, int device, uint32_t rowid_base, mgpu_index ** out )
{
	if ( !path_prefix || !out )
		return MGPU_E_BAD_QUERY;
	*out = nullptr;
	std::unique_ptr<mgpu_index> p ( new mgpu_index );
	int iRes = p->m_t.Open ( path_prefix, device, rowid_base );
	if ( iRes!=MGPU_OK )
	{
		g_sLastOpenError = p->m_t.m_sError;
		return iRes;
	}
	*out = p.release();
	return MGPU_OK;
}

int mgpu_index_close ( mgpu_index * idx )
{
	if ( !idx )
		return MGPU_OK;
	{
		// batches hold a plain pointer to their index: closing under them is refused (the handle stays valid), and whatever the
		// handle's streams still run is drained before the device memory goes away
		std::lock_guard<std::mutex> tGuard ( idx->m_t.m_tLock );
		if ( idx->m_t.m_nLiveBatches.load()>0 )
		{
			idx->m_t.m_sError = "mgpu_index_close: batches of this index are still alive (free them first)";
			return MGPU_E_BAD_QUERY;
		}
		cudaSetDevice ( idx->m_t.m_iDevice );
		if ( idx->m_t.m_tOwnStream )
			cudaStreamSynchronize ( idx->m_t.m_tOwnStream );
		if ( idx->m_t.m_tHotStream )
			cudaStreamSynchronize ( idx->m_t.m_tHotStream );
	}
	delete idx;
	return MGPU_OK;
}

int mgpu_index_set_stream ( mgpu_index * idx, void * cuda_stream )
{
	if ( !idx )
		return MGPU_E_BAD_QUERY;
	std::lock_guard<std::mutex> tGuard ( idx->m_t.m_tLock );
	cudaSetDevice ( idx->m_t.m_iDevice );
	cudaStreamSynchronize ( idx->m_t.m_tStream );
	idx->m_t.m_tStream = cuda_stream ? (cudaStream_t)cuda_stream : idx->m_t.m_tOwnStream;
	return MGPU_OK;
}

int mgpu_index_set_option ( mgpu_index * idx, const char * name, int64_t value )
{
	if ( !idx || !name )
		return MGPU_E_BAD_QUERY;
	std::lock_guard<std::mutex> tGuard ( idx->m_t.m_tLock );
	if ( !idx->m_t.m_tOpt.Set ( name, value ) )
	{
		idx->m_t.m_sError = std::string ( "unknown option or value out of range: " ) + name;
		return MGPU_E_BAD_QUERY;
	}
	return MGPU_OK;
}

const char * mgpu_last_error ( const mgpu_index * idx )
{
	return idx ? idx->m_t.m_sError.c_str() : g_sLastOpenError.c_str();
}

int64_t mgpu_index_total_docs ( const mgpu_index * idx )	{ return idx ? (int64_t)idx->m_t.m_tHdr.m_iTotalDocuments : 0; }
int32_t mgpu_index_num_fields ( const mgpu_index * idx )	{ return idx ? (int32_t)idx->m_t.m_tHdr.m_dFields.size() : 0; }
int32_t mgpu_index_field_index ( const mgpu_index * idx, const char * name )	{ return ( idx && name ) ? idx->m_t.FieldIndex ( name ) : -1; }
const char * mgpu_index_field_name ( const mgpu_index * idx, int32_t field )
{
	return ( idx && field>=0 && field<(int32_t)idx->m_t.m_tHdr.m_dFields.size() ) ? idx->m_t.m_tHdr.m_dFields[field].m_sName.c_str() : nullptr;
}
int32_t mgpu_index_attr_index ( const mgpu_index * idx, const char * name )	{ return ( idx && name ) ? idx->m_t.AttrIndex ( name ) : -1; }

int mgpu_index_word_stats ( const mgpu_index * idx, const char * word, int64_t * docs, int64_t * hits )
{
	const TermInfo_t * p = ( idx && word ) ? idx->m_t.FindTerm ( word ) : nullptr;
	if ( !p )
		return 0;
	if ( docs ) *docs = p->m_iDocs;
	if ( hits ) *hits = p->m_iHits;
	return 1;
}

int mgpu_index_word_bytes ( const mgpu_index * idx, const char * word, int64_t * doclist_bytes, int64_t * skiplist_bytes )
{
	const TermInfo_t * p = ( idx && word ) ? idx->m_t.FindTerm ( word ) : nullptr;
	if ( !p )
		return 0;
	if ( doclist_bytes ) *doclist_bytes = p->m_iDoclistLength;
	if ( skiplist_bytes ) *skiplist_bytes = p->m_iSkiplistBytes;
	return 1;
}

int mgpu_batch_prepare ( mgpu_index * idx, const mgpu_query * queries, int n_queries, mgpu_batch ** out )
{
	if ( !idx || !out || n_queries<0 || ( n_queries && !queries ) )
		return MGPU_E_BAD_QUERY;
	*out = nullptr;
	std::unique_ptr<mgpu_batch> p ( new mgpu_batch );
	std::lock_guard<std::mutex> tGuard ( idx->m_t.m_tLock );
	int iRes = p->m_t.Prepare ( &idx->m_t, queries, n_queries );
	if ( iRes!=MGPU_OK )
	{
		idx->m_t.m_sError = p->m_t.m_sError;
		return iRes;
	}
	*out = p.release();
	return MGPU_OK;
}

int mgpu_batch_run ( mgpu_batch * b )
{
	if ( !b )
		return MGPU_E_BAD_QUERY;
	std::lock_guard<std::mutex> tGuard ( b->m_t.m_pIndex->m_tLock );
	int iRes = b->m_t.Run();
	if ( iRes!=MGPU_OK )
		b->m_t.m_pIndex->m_sError = b->m_t.m_sError;
	return iRes;
}

int mgpu_batch_sync ( mgpu_batch * b )
{
	if ( !b )
		return MGPU_E_BAD_QUERY;
	int iRes = b->m_t.Sync();
	if ( iRes!=MGPU_OK )
		b->m_t.m_pIndex->m_sError = b->m_t.m_sError;
	return iRes;
}

int mgpu_batch_fetch ( mgpu_batch * b, mgpu_result * results )
{
	if ( !b || !results )
		return MGPU_E_BAD_QUERY;
	int iRes = b->m_t.Fetch ( results );
	if ( iRes!=MGPU_OK )
		b->m_t.m_pIndex->m_sError = b->m_t.m_sError;
	return iRes;
}

void mgpu_batch_free ( mgpu_batch * b )
{
	delete b;
}

int mgpu_batch_get_stats ( const mgpu_batch * b, mgpu_batch_stats * out )
{
	if ( !b || !out )
		return MGPU_E_BAD_QUERY;
	*out = b->m_t.m_tStats;
	return MGPU_OK;
}

int mgpu_search_batch ( mgpu_index * idx, const mgpu_query * queries, int n_queries, mgpu_result * results )
{
	const auto tStart = std::chrono::steady_clock::now();
	if ( !idx || n_queries<0 || ( n_queries && ( !queries || !results ) ) )
		return MGPU_E_BAD_QUERY;
	std::unique_ptr<mgpu_batch> pBatch ( new mgpu_batch );
	mgpu_batch * b = pBatch.get();
	int iRes;
	{
		// prepare + run under ONE lock: no other batch of this handle can get between them, so the hot-term store may be
		// started from inside Prepare (it then overlaps with the rest of the host-side setup)
		std::lock_guard<std::mutex> tGuard ( idx->m_t.m_tLock );
		iRes = b->m_t.Prepare ( &idx->m_t, queries, n_queries, nullptr, 0, idx->m_t.m_tOpt.m_bEagerHot!=0 );
		if ( iRes==MGPU_OK )
			iRes = b->m_t.Run();
		if ( iRes!=MGPU_OK )
		{
			idx->m_t.m_sError = b->m_t.m_sError;
			return iRes;
		}
	}
	iRes = mgpu_batch_fetch ( b, results );
	mgpu_batch_stats tStats = b->m_t.m_tStats;
	pBatch.reset();
	tStats.host_total_ms = std::chrono::duration<float,std::milli> ( std::chrono::steady_clock::now()-tStart ).count();
	{
		std::lock_guard<std::mutex> tGuard ( idx->m_t.m_tLock );
		idx->m_t.m_tLastSearchStats = tStats;
	}
	return iRes;
}

int mgpu_index_last_search_stats ( const mgpu_index * idx, mgpu_batch_stats * out )
{
	if ( !idx || !out )
		return MGPU_E_BAD_QUERY;
	*out = idx->m_t.m_tLastSearchStats;
	return MGPU_OK;
}

int mgpu_batch_export_keys ( mgpu_batch * b, void * dev_keys, void * dev_counts, void * dev_total_found, int K )
{
	if ( !b || !dev_keys || !dev_counts || !dev_total_found || K<1 )
		return MGPU_E_BAD_QUERY;
	int iRes = b->m_t.ExportKeys ( dev_keys, dev_counts, dev_total_found, K );	// asynchronous, ordered on the index stream
	if ( iRes!=MGPU_OK )
		b->m_t.m_pIndex->m_sError = b->m_t.m_sError;
	return iRes;
}

int mgpu_merge_shard_keys ( int device, const void * dev_keys, const void * dev_counts, int n_shards, int nq, int K,
	void * dev_out_keys, void * dev_out_counts, void * stream )
{
	if ( !dev_keys || !dev_counts || !dev_out_keys || !dev_out_counts || n_shards<1 || nq<0 || K<1 )
		return MGPU_E_BAD_QUERY;
	if ( !nq )
		return MGPU_OK;
	if ( cudaSetDevice ( device )!=cudaSuccess )
		return MGPU_E_NO_DEVICE;
	int iStride = 2;
	while ( iStride<2*n_shards*K )
		iStride <<= 1;
	// stream-ordered scratch from the device pool: asynchronous, no device-wide synchronisation (the caller orders on `stream`)
	cudaStream_t s = (cudaStream_t)stream;
	Key128_t * pScratch = nullptr;
	if ( cudaMallocAsync ( (void**)&pScratch, (size_t)nq*iStride*sizeof(Key128_t), s )!=cudaSuccess )
		return MGPU_E_NOMEM;
	cudaError_t e = LaunchShardMerge ( (const Key128_t*)dev_keys, (const int32_t*)dev_counts, n_shards, nq, K, pScratch, iStride,
		(Key128_t*)dev_out_keys, (int32_t*)dev_out_counts, nq<1184 ? nq : 1184, s );
	cudaFreeAsync ( pScratch, s );
	return e==cudaSuccess ? MGPU_OK : MGPU_E_CUDA;
}

void mgpu_unpack_key ( const uint64_t key[2], uint32_t * global_rowid, int32_t * weight, uint64_t * sortkey_hi )
{
	if ( global_rowid ) *global_rowid = ~(uint32_t)( key[1]>>32 );
	if ( weight ) *weight = (int32_t)(uint32_t)key[1];
	if ( sortkey_hi ) *sortkey_hi = key[0];
}

int mgpu_decode_doclist ( mgpu_index * idx, const char * word, uint32_t * rowid, uint32_t * hits, uint32_t * fields, uint64_t * hitlist_pos, int64_t capacity, int64_t * n_out )
{
	if ( !idx || !word || !n_out )
		return MGPU_E_BAD_QUERY;
	Index_c & t = idx->m_t;
	*n_out = 0;
	const TermInfo_t * pTerm = t.FindTerm ( word );
	if ( !pTerm )
		return MGPU_OK;
	*n_out = pTerm->m_iDocs;
	std::lock_guard<std::mutex> tGuard ( t.m_tLock );
	if ( cudaSetDevice ( t.m_iDevice )!=cudaSuccess )
		return MGPU_E_NO_DEVICE;

	DevLeaf_t tLeaf {};
	tLeaf.m_uFirstBlk = pTerm->m_uFirstBlk;
	tLeaf.m_nBlocks = pTerm->m_nBlocks;
	tLeaf.m_nDocs = (uint32_t)pTerm->m_iDocs;
	tLeaf.m_uDoclistEnd = (uint64_t)( pTerm->m_iDoclistOffset+pTerm->m_iDoclistLength-1 );
	tLeaf.m_uQueriedFields = 0xFFFFFFFFu;

	const size_t n = (size_t)pTerm->m_nBlocks*32;
	const bool bStore = ( rowid && hits && fields && hitlist_pos && capacity>=pTerm->m_iDocs );
	DevBuf_T<uint32_t> dRow, dHits, dFields;
	DevBuf_T<uint64_t> dPos;
	DevBuf_T<unsigned long long> dSum;
	if ( dSum.Alloc ( 1 )!=cudaSuccess )
		return MGPU_E_NOMEM;
	cudaMemsetAsync ( dSum.m_p, 0, 8, t.m_tStream );
	if ( bStore && ( dRow.Alloc ( n )!=cudaSuccess || dHits.Alloc ( n )!=cudaSuccess || dFields.Alloc ( n )!=cudaSuccess || dPos.Alloc ( n )!=cudaSuccess ) )
		return MGPU_E_NOMEM;
	int nCtas = (int)std::min<size_t> ( ( pTerm->m_nBlocks+EVAL_WARPS-1 )/EVAL_WARPS, (size_t)t.m_nSMs*8 );
	cudaError_t e = LaunchDecodeDoclist ( t.m_tDev, tLeaf, dRow.m_p, dHits.m_p, dFields.m_p, dPos.m_p, dSum.m_p, nCtas, t.m_tStream );
	if ( e==cudaSuccess && bStore )
	{
		const size_t nDocs = (size_t)pTerm->m_iDocs;
		cudaMemcpyAsync ( rowid, dRow.m_p, nDocs*4, cudaMemcpyDeviceToHost, t.m_tStream );
		cudaMemcpyAsync ( hits, dHits.m_p, nDocs*4, cudaMemcpyDeviceToHost, t.m_tStream );
		cudaMemcpyAsync ( fields, dFields.m_p, nDocs*4, cudaMemcpyDeviceToHost, t.m_tStream );
		cudaMemcpyAsync ( hitlist_pos, dPos.m_p, nDocs*8, cudaMemcpyDeviceToHost, t.m_tStream );
	}
	if ( e==cudaSuccess )
		e = cudaStreamSynchronize ( t.m_tStream );
	if ( e!=cudaSuccess )
	{
		t.m_sError = std::string ( "decode_doclist: " ) + cudaGetErrorString ( e );
		return MGPU_E_CUDA;
	}
	return MGPU_OK;
}

} // extern "C"
