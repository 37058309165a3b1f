// Rowid-range shards of one index behind one handle: the reference's "distributed index of local indexes"
// (RunLocalSearches, src/searchd.cpp:5596-5814: one thread per local index, every index with its own sorter;
// MergeAllMatches, :4653-4738; global IDF via CSphMultiQueryArgs::m_iTotalDocs / m_pLocalDocs filled by SetupLocalDF, :5869)
// with one GPU per shard. One process, one host thread per shard:
//
//   1. the batch is planned ONCE (against shard 0's dictionary, all host threads) with the statistics of the whole index
//      (total documents, global df per keyword: summed over the shards' dictionaries when the handle is opened);
//   2. every shard thread binds that plan to its own dictionary (Batch_c::Prepare with a template), uploads it, runs the batch
//      on its GPU and exports its local top-K keys + document ids;
//   3. the K keys per query and shard travel to shard 0's GPU - ncclSend / ncclRecv over NVLink when the shards sit on
//      different GPUs (NCCL is loaded at run time; the single-index library does not depend on it), plain stream-ordered
//      device copies when they share one - and shard_merge_kernel keeps the best K (disjoint rowid ranges: nothing to dedupe,
//      unlike KillPlainDupes, src/searchd.cpp:3910-3952); total_found and the keyword statistics are summed on the host;
//   4. one download through pinned memory fills the caller's host buffers.
//
// Results are those of the UNSHARDED index: global N / df for the IDFs, keyword order of multi-keyword nodes by global df
// (mgpu_query::shard_of_global), ties broken by the global rowid packed into the keys.
#include "engine.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>

#include <dlfcn.h>

namespace mgpu
{

#define SH_TRY(_expr) \
	do { cudaError_t _e = (_expr); if ( _e!=cudaSuccess ) { m_sError = std::string ( #_expr ": " ) + cudaGetErrorString ( _e ); return MGPU_E_CUDA; } } while (0)

// the few NCCL entry points the exchange needs, resolved from libnccl.so.2 at run time
namespace
{
typedef void * NcclComm_t;
struct NcclApi_t
{
	void *	m_pLib = nullptr;
	int		( *CommInitAll ) ( NcclComm_t *, int, const int * ) = nullptr;
	int		( *CommDestroy ) ( NcclComm_t ) = nullptr;
	int		( *GroupStart ) () = nullptr;
	int		( *GroupEnd ) () = nullptr;
	int		( *Send ) ( const void *, size_t, int, int, NcclComm_t, cudaStream_t ) = nullptr;
	int		( *Recv ) ( void *, size_t, int, int, NcclComm_t, cudaStream_t ) = nullptr;
	const char * ( *GetErrorString ) ( int ) = nullptr;
	bool Load ( std::string & sError )
	{
		if ( m_pLib )
			return true;
		const char * dNames[] = { "libnccl.so.2", "libnccl.so" };
		for ( const char * szName : dNames )
			if ( ( m_pLib = dlopen ( szName, RTLD_NOW | RTLD_GLOBAL ) )!=nullptr )
				break;
		if ( !m_pLib )
		{
			sError = std::string ( "cannot load NCCL: " ) + dlerror();
			return false;
		}
		CommInitAll = ( decltype(CommInitAll) ) dlsym ( m_pLib, "ncclCommInitAll" );
		CommDestroy = ( decltype(CommDestroy) ) dlsym ( m_pLib, "ncclCommDestroy" );
		GroupStart = ( decltype(GroupStart) ) dlsym ( m_pLib, "ncclGroupStart" );
		GroupEnd = ( decltype(GroupEnd) ) dlsym ( m_pLib, "ncclGroupEnd" );
		Send = ( decltype(Send) ) dlsym ( m_pLib, "ncclSend" );
		Recv = ( decltype(Recv) ) dlsym ( m_pLib, "ncclRecv" );
		GetErrorString = ( decltype(GetErrorString) ) dlsym ( m_pLib, "ncclGetErrorString" );
		if ( !CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !Send || !Recv || !GetErrorString )
		{
			sError = "libnccl lacks a symbol the shard exchange needs";
			return false;
		}
		return true;
	}
};
const int NCCL_UINT8 = 1;	// ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)
}

class ShardedIndex_c
{
public:
	std::string										m_sError;
	std::vector<std::unique_ptr<Index_c>>			m_dShards;
	std::vector<uint32_t>							m_dBase;		///< [shards+1] global rowid of every shard's row 0, and the total
	struct GlobalWord_t { int64_t m_iDocs = 0, m_iHits = 0; int32_t m_iId = -1; };
	std::unordered_map<std::string,GlobalWord_t>	m_hGlobalDocs;	///< keyword -> documents / hits over all shards, id in the tables below
	std::vector<std::vector<const TermInfo_t*>>		m_dTermOfId;	///< [shard][keyword id] -> the shard's dictionary entry (null: not in this shard)
	int64_t											m_iTotalDocs = 0;
	std::mutex										m_tLock;		///< one batch at a time per handle
	mgpu_sharded_stats								m_tStats {};

	NcclApi_t										m_tNccl;
	std::vector<NcclComm_t>							m_dComms;		///< one per shard when every shard has a GPU of its own
	bool											m_bSameDevice = true;

	// host arrays reused from call to call (their pages stay mapped)
	std::vector<mgpu_query>			m_dQueries;
	std::vector<int64_t>			m_dWordDocs, m_dWordHits;
	std::vector<int32_t>			m_dWordIds;
	std::vector<PlannedQuery_t>		m_dTemplate;
	std::vector<DevQueryExt_t>		m_dTemplateExt;

	// exchange buffers on shard 0's GPU (grow-only)
	DevBuf_T<Key128_t>		m_dGatherKeys, m_dMergeScratch, m_dOutKeys;
	DevBuf_T<int64_t>		m_dGatherDocid, m_dOutDocid, m_dGatherTotal;
	DevBuf_T<int32_t>		m_dGatherCount, m_dOutCount;
	DevBuf_T<uint32_t>		m_dBaseDev;
	// per-shard send buffers (shards on other GPUs)
	struct Send_t { DevBuf_T<Key128_t> m_dKeys; DevBuf_T<int64_t> m_dDocid, m_dTotal; DevBuf_T<int32_t> m_dCount; };
	std::vector<std::unique_ptr<Send_t>>	m_dSend;
	std::vector<cudaEvent_t>				m_dEvents;

	~ShardedIndex_c()
	{
		for ( NcclComm_t c : m_dComms )
			if ( c )
				m_tNccl.CommDestroy ( c );
		for ( size_t s=0; s<m_dEvents.size(); ++s )
			if ( m_dEvents[s] )
			{
				cudaSetDevice ( m_dShards[s]->m_iDevice );
				cudaEventDestroy ( m_dEvents[s] );
			}
		if ( !m_dShards.empty() )
			cudaSetDevice ( m_dShards[0]->m_iDevice );	// the exchange buffers live there
	}

	int Open ( const char * const * dPrefixes, const int * dDevices, int nShards )
	{
		uint64_t uBase = 0;
		for ( int s=0; s<nShards; ++s )
		{
			std::unique_ptr<Index_c> p ( new Index_c );
			if ( uBase>0xFFFFFFFFull )
			{
				m_sError = "more than 2^32 rows over all shards";
				return MGPU_E_UNSUPPORTED;
			}
			int iRes = p->Open ( dPrefixes[s], dDevices[s], (uint32_t)uBase );
			if ( iRes!=MGPU_OK )
			{
				m_sError = p->m_sError;
				return iRes;
			}
			if ( s && p->m_tHdr.m_bWordDict!=m_dShards[0]->m_tHdr.m_bWordDict )
			{
				m_sError = "shards mix dict=keywords and dict=crc";
				return MGPU_E_FORMAT;
			}
			m_dBase.push_back ( (uint32_t)uBase );
			uBase += p->m_tDev.m_uRows;
			m_iTotalDocs += (int64_t)p->m_tHdr.m_iTotalDocuments;
			for ( const auto & kv : p->m_hTerms )
			{
				GlobalWord_t & t = m_hGlobalDocs[kv.first];
				if ( t.m_iId<0 )
					t.m_iId = (int32_t)m_hGlobalDocs.size()-1;
				t.m_iDocs += kv.second.m_iDocs;
				t.m_iHits += kv.second.m_iHits;
			}
			m_bSameDevice = m_bSameDevice && dDevices[s]==dDevices[0];
			m_dShards.push_back ( std::move ( p ) );
		}
		m_dBase.push_back ( (uint32_t)std::min<uint64_t> ( uBase, 0xFFFFFFFFull ) );
		// every shard's dictionary entry per global keyword id: the shard threads bind a batch's keywords by id, not by name
		m_dTermOfId.assign ( nShards, std::vector<const TermInfo_t*> ( m_hGlobalDocs.size(), nullptr ) );
		for ( int s=0; s<nShards; ++s )
			for ( const auto & kv : m_dShards[s]->m_hTerms )
				m_dTermOfId[s][m_hGlobalDocs[kv.first].m_iId] = &kv.second;
		bool bDistinct = true;
		for ( int s=0; s<nShards; ++s )
			for ( int t=0; t<s; ++t )
				bDistinct = bDistinct && dDevices[s]!=dDevices[t];
		if ( nShards>1 && bDistinct )
		{
			if ( !m_tNccl.Load ( m_sError ) )
				return MGPU_E_IO;
			m_dComms.assign ( nShards, nullptr );
			int iRes = m_tNccl.CommInitAll ( m_dComms.data(), nShards, dDevices );
			if ( iRes!=0 )
			{
				m_sError = std::string ( "ncclCommInitAll: " ) + m_tNccl.GetErrorString ( iRes );
				m_dComms.clear();
				return MGPU_E_CUDA;
			}
		}
		m_dEvents.assign ( nShards, nullptr );
		for ( int s=0; s<nShards; ++s )
		{
			SH_TRY ( cudaSetDevice ( dDevices[s] ) );
			SH_TRY ( cudaEventCreateWithFlags ( &m_dEvents[s], cudaEventDisableTiming ) );
			m_dSend.emplace_back ( new Send_t );
		}
		SH_TRY ( cudaSetDevice ( dDevices[0] ) );
		SH_TRY ( m_dBaseDev.Alloc ( m_dBase.size() ) );
		SH_TRY ( cudaMemcpy ( m_dBaseDev.m_p, m_dBase.data(), m_dBase.size()*4, cudaMemcpyHostToDevice ) );
		return MGPU_OK;
	}

	int Search ( const mgpu_query * pQueries, int nQueries, mgpu_result * pResults )
	{
		const auto tStart = std::chrono::steady_clock::now();
		auto fnMs = [] ( std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b ) { return std::chrono::duration<float,std::milli> ( b-a ).count(); };
		const int nShards = (int)m_dShards.size();
		m_tStats = mgpu_sharded_stats {};
		m_tStats.n_shards = nShards;

		// 1. statistics of the whole index for every query that does not bring its own (SetupLocalDF, src/searchd.cpp:5869) and
		// 2. planning, ONCE (shard 0's dictionary; the shard threads re-bind the keywords): both per query, spread over the host threads.
		// The arrays are members: a 10k-query batch's 34 MB of plans are not page-faulted in again on every call.
		std::vector<mgpu_query> & dQueries = m_dQueries;
		dQueries.assign ( pQueries, pQueries+nQueries );
		std::vector<size_t> dWordOff ( nQueries+1, 0 );
		for ( int i=0; i<nQueries; ++i )
			dWordOff[i+1] = dWordOff[i] + (size_t)std::max ( dQueries[i].n_words, 0 );
		std::vector<int64_t> & dWordDocs = m_dWordDocs;
		dWordDocs.resize ( dWordOff[nQueries] );
		std::vector<int32_t> & dWordIds = m_dWordIds;
		dWordIds.resize ( dWordOff[nQueries] );
		std::vector<int64_t> & dWordHits = m_dWordHits;
		dWordHits.resize ( dWordOff[nQueries] );
		std::vector<PlannedQuery_t> & dTemplate = m_dTemplate;
		dTemplate.clear();
		dTemplate.resize ( nQueries );
		{
			int nThreads = (int)std::min<unsigned> ( std::max ( 1u, std::thread::hardware_concurrency() ), 32u );
			nThreads = std::max ( 1, std::min ( nThreads, nQueries/256 ) );
			if ( m_dShards[0]->m_tOpt.m_iPlanThreads>0 )
				nThreads = m_dShards[0]->m_tOpt.m_iPlanThreads;
			std::vector<std::vector<std::pair<int,DevQueryExt_t>>> dThreadExt ( std::max ( nThreads, 1 ) );
			auto fnPlan = [&] ( int iFrom, int iTo, int iThread )
			{
				DevQueryExt_t tExt;
				for ( int i=iFrom; i<iTo; ++i )
				{
					mgpu_query & q = dQueries[i];
					if ( !q.total_docs )
						q.total_docs = m_iTotalDocs;
					if ( q.words && q.n_words>0 )
					{
						// one lookup per keyword for the whole handle: statistics of the whole index, and the id the shards bind it by
						int64_t * pDocs = dWordDocs.data()+dWordOff[i], * pHits = dWordHits.data()+dWordOff[i];
						int32_t * pIds = dWordIds.data()+dWordOff[i];
						for ( int w=0; w<q.n_words; ++w )
						{
							pDocs[w] = -1;
							pHits[w] = 0;
							pIds[w] = -1;
							if ( q.words[w].word )
							{
								auto it = m_hGlobalDocs.find ( m_dShards[0]->DictKey ( q.words[w].word ) );
								pDocs[w] = it==m_hGlobalDocs.end() ? 0 : it->second.m_iDocs;
								pHits[w] = it==m_hGlobalDocs.end() ? 0 : it->second.m_iHits;
								pIds[w] = it==m_hGlobalDocs.end() ? -1 : it->second.m_iId;
							}
						}
						if ( !q.word_docs )
							q.word_docs = pDocs;
					}
					q.shard_of_global = 1;
					bool bHasExt = false;
					PlanQuery ( *m_dShards[0], q, dTemplate[i], tExt, bHasExt );
					if ( bHasExt )
						dThreadExt[iThread].push_back ( { i, tExt } );
					// the keywords' statistics reported with the result: the whole index's (the shards' dictionaries summed at open)
					for ( size_t w=0; w<dTemplate[i].m_dWordStats.size(); ++w )
						dTemplate[i].m_dWordStats[w] = mgpu_wordstat { std::max<int64_t> ( dWordDocs[dWordOff[i]+w], 0 ), dWordHits[dWordOff[i]+w] };
				}
			};
			if ( nThreads<=1 )
				fnPlan ( 0, nQueries, 0 );
			else
			{
				std::vector<std::thread> dThreads;
				for ( int t=0; t<nThreads; ++t )
					dThreads.emplace_back ( fnPlan, (int)( (int64_t)nQueries*t/nThreads ), (int)( (int64_t)nQueries*( t+1 )/nThreads ), t );
				for ( auto & t : dThreads )
					t.join();
			}
			// the queries' extensions (filters, sort keys, hit-level nodes), numbered in query order; every shard uploads this array
			m_dTemplateExt.clear();
			for ( const auto & dMine : dThreadExt )
				for ( const auto & t : dMine )
				{
					dTemplate[t.first].m_tDev.m_iExt = (int)m_dTemplateExt.size();
					m_dTemplateExt.push_back ( t.second );
				}
		}
		const auto tPlanned = std::chrono::steady_clock::now();
		m_tStats.host_plan_ms = fnMs ( tStart, tPlanned );

		// 3. one thread per shard: bind + upload + run + export (asynchronous on the shard's stream)
		int iKMax = 1;
		for ( int i=0; i<nQueries; ++i )
			if ( dTemplate[i].m_iStatus==MGPU_OK && dTemplate[i].m_tDev.m_nOps>0 )
				iKMax = std::max ( iKMax, dTemplate[i].m_tDev.m_iMaxMatches );
		const size_t nSlots = (size_t)nQueries*iKMax;
		const int iRootDev = m_dShards[0]->m_iDevice;
		SH_TRY ( cudaSetDevice ( iRootDev ) );
		SH_TRY ( m_dGatherKeys.Grow ( nSlots*nShards ) );
		SH_TRY ( m_dGatherDocid.Grow ( nSlots*nShards ) );
		SH_TRY ( m_dGatherCount.Grow ( (size_t)nQueries*nShards ) );
		SH_TRY ( m_dGatherTotal.Grow ( (size_t)nQueries*nShards ) );
		SH_TRY ( m_dOutKeys.Grow ( nSlots ) );
		SH_TRY ( m_dOutDocid.Grow ( nSlots ) );
		SH_TRY ( m_dOutCount.Grow ( nQueries ) );
		int iScratchStride = 2;
		while ( iScratchStride<2*nShards*iKMax )
			iScratchStride <<= 1;
		SH_TRY ( m_dMergeScratch.Grow ( (size_t)nQueries*iScratchStride ) );

		std::vector<std::unique_ptr<Batch_c>> dBatches ( nShards );
		std::vector<int> dRes ( nShards, MGPU_OK );
		const int nBindThreads = std::max ( 1, (int)std::thread::hardware_concurrency()/std::max ( 1, nShards ) );
		auto fnShard = [&] ( int s )
		{
			Index_c * pIndex = m_dShards[s].get();
			dBatches[s].reset ( new Batch_c );
			Batch_c & b = *dBatches[s];
			int iRes = b.Prepare ( pIndex, dQueries.data(), nQueries, &dTemplate, nBindThreads, pIndex->m_tOpt.m_bEagerHot!=0,
				m_dWordIds.data(), dWordOff.data(), m_dTermOfId[s].data(), &m_dTemplateExt );
			if ( iRes==MGPU_OK )
				iRes = b.Run();
			if ( iRes==MGPU_OK )
			{
				const bool bLocal = pIndex->m_iDevice==iRootDev;
				Send_t & tSend = *m_dSend[s];
				void * pKeys, * pDocid, * pCount, * pTotal;
				if ( bLocal )
				{
					// same GPU as the merge: export straight into the gather buffers
					pKeys = m_dGatherKeys.m_p + nSlots*s;
					pDocid = m_dGatherDocid.m_p + nSlots*s;
					pCount = m_dGatherCount.m_p + (size_t)nQueries*s;
					pTotal = m_dGatherTotal.m_p + (size_t)nQueries*s;
				} else
				{
					cudaSetDevice ( pIndex->m_iDevice );
					if ( tSend.m_dKeys.Grow ( nSlots )!=cudaSuccess || tSend.m_dDocid.Grow ( nSlots )!=cudaSuccess
						|| tSend.m_dCount.Grow ( nQueries )!=cudaSuccess || tSend.m_dTotal.Grow ( nQueries )!=cudaSuccess )
						iRes = MGPU_E_NOMEM;
					pKeys = tSend.m_dKeys.m_p; pDocid = tSend.m_dDocid.m_p; pCount = tSend.m_dCount.m_p; pTotal = tSend.m_dTotal.m_p;
				}
				if ( iRes==MGPU_OK )
					iRes = b.ExportKeys ( pKeys, pCount, pTotal, iKMax, pDocid );
				if ( iRes==MGPU_OK && cudaEventRecord ( m_dEvents[s], b.m_tStream )!=cudaSuccess )
					iRes = MGPU_E_CUDA;
			}
			dRes[s] = iRes;
		};
		if ( nShards==1 )
			fnShard ( 0 );
		else
		{
			std::vector<std::thread> dThreads;
			for ( int s=0; s<nShards; ++s )
				dThreads.emplace_back ( fnShard, s );
			for ( auto & t : dThreads )
				t.join();
		}
		for ( int s=0; s<nShards; ++s )
			if ( dRes[s]!=MGPU_OK )
			{
				m_sError = dBatches[s] ? dBatches[s]->m_sError : "shard thread failed";
				return dRes[s];
			}
		const auto tLaunched = std::chrono::steady_clock::now();
		m_tStats.host_setup_ms = fnMs ( tPlanned, tLaunched );

		// 4. exchange: every shard's keys to shard 0's GPU, then the merge there
		SH_TRY ( cudaSetDevice ( iRootDev ) );
		cudaStream_t tRoot = m_dShards[0]->m_tStream;
		if ( !m_dComms.empty() )
		{
			// ncclSend on the shard's stream (ordered after its export), ncclRecv on the root's stream
			int iRes = m_tNccl.GroupStart();
			for ( int s=1; s<nShards && iRes==0; ++s )
			{
				Send_t & tSend = *m_dSend[s];
				cudaStream_t tStream = m_dShards[s]->m_tStream;
				iRes = m_tNccl.Send ( tSend.m_dKeys.m_p, nSlots*sizeof(Key128_t), NCCL_UINT8, 0, m_dComms[s], tStream );
				if ( !iRes ) iRes = m_tNccl.Send ( tSend.m_dDocid.m_p, nSlots*8, NCCL_UINT8, 0, m_dComms[s], tStream );
				if ( !iRes ) iRes = m_tNccl.Send ( tSend.m_dCount.m_p, (size_t)nQueries*4, NCCL_UINT8, 0, m_dComms[s], tStream );
				if ( !iRes ) iRes = m_tNccl.Send ( tSend.m_dTotal.m_p, (size_t)nQueries*8, NCCL_UINT8, 0, m_dComms[s], tStream );
				if ( !iRes ) iRes = m_tNccl.Recv ( m_dGatherKeys.m_p + nSlots*s, nSlots*sizeof(Key128_t), NCCL_UINT8, s, m_dComms[0], tRoot );
				if ( !iRes ) iRes = m_tNccl.Recv ( m_dGatherDocid.m_p + nSlots*s, nSlots*8, NCCL_UINT8, s, m_dComms[0], tRoot );
				if ( !iRes ) iRes = m_tNccl.Recv ( m_dGatherCount.m_p + (size_t)nQueries*s, (size_t)nQueries*4, NCCL_UINT8, s, m_dComms[0], tRoot );
				if ( !iRes ) iRes = m_tNccl.Recv ( m_dGatherTotal.m_p + (size_t)nQueries*s, (size_t)nQueries*8, NCCL_UINT8, s, m_dComms[0], tRoot );
			}
			const int iEnd = m_tNccl.GroupEnd();
			if ( iRes || iEnd )
			{
				m_sError = std::string ( "NCCL exchange: " ) + m_tNccl.GetErrorString ( iRes ? iRes : iEnd );
				return MGPU_E_CUDA;
			}
			m_tStats.nccl = 1;
		} else
			for ( int s=1; s<nShards; ++s )
			{
				SH_TRY ( cudaStreamWaitEvent ( tRoot, m_dEvents[s], 0 ) );
				if ( m_dShards[s]->m_iDevice!=iRootDev )
				{
					// (shards sharing GPUs unevenly: no communicator; peer copies ordered on the root's stream)
					Send_t & tSend = *m_dSend[s];
					const int iDev = m_dShards[s]->m_iDevice;
					SH_TRY ( cudaMemcpyPeerAsync ( m_dGatherKeys.m_p + nSlots*s, iRootDev, tSend.m_dKeys.m_p, iDev, nSlots*sizeof(Key128_t), tRoot ) );
					SH_TRY ( cudaMemcpyPeerAsync ( m_dGatherDocid.m_p + nSlots*s, iRootDev, tSend.m_dDocid.m_p, iDev, nSlots*8, tRoot ) );
					SH_TRY ( cudaMemcpyPeerAsync ( m_dGatherCount.m_p + (size_t)nQueries*s, iRootDev, tSend.m_dCount.m_p, iDev, (size_t)nQueries*4, tRoot ) );
					SH_TRY ( cudaMemcpyPeerAsync ( m_dGatherTotal.m_p + (size_t)nQueries*s, iRootDev, tSend.m_dTotal.m_p, iDev, (size_t)nQueries*8, tRoot ) );
				}
			}
		SH_TRY ( LaunchShardMerge ( m_dGatherKeys.m_p, m_dGatherCount.m_p, nShards, nQueries, iKMax, m_dMergeScratch.m_p, iScratchStride,
			m_dOutKeys.m_p, m_dOutCount.m_p, std::max ( 1, std::min ( nQueries, m_dShards[0]->m_nSMs*8 ) ), tRoot, m_dGatherDocid.m_p, m_dBaseDev.m_p, m_dOutDocid.m_p ) );

		// 5. one download through the root's pinned staging
		const size_t iOffDocid = nSlots*sizeof(Key128_t), iOffTotal = iOffDocid + nSlots*8, iOffCount = iOffTotal + (size_t)nQueries*nShards*8;
		uint8_t * pStage = (uint8_t *)m_dShards[0]->Pinned ( iOffCount + (size_t)nQueries*4 );
		if ( !pStage )
		{
			m_sError = "cudaHostAlloc failed";
			return MGPU_E_NOMEM;
		}
		SH_TRY ( cudaMemcpyAsync ( pStage, m_dOutKeys.m_p, nSlots*sizeof(Key128_t), cudaMemcpyDeviceToHost, tRoot ) );
		SH_TRY ( cudaMemcpyAsync ( pStage+iOffDocid, m_dOutDocid.m_p, nSlots*8, cudaMemcpyDeviceToHost, tRoot ) );
		SH_TRY ( cudaMemcpyAsync ( pStage+iOffTotal, m_dGatherTotal.m_p, (size_t)nQueries*nShards*8, cudaMemcpyDeviceToHost, tRoot ) );
		SH_TRY ( cudaMemcpyAsync ( pStage+iOffCount, m_dOutCount.m_p, (size_t)nQueries*4, cudaMemcpyDeviceToHost, tRoot ) );
		SH_TRY ( cudaStreamSynchronize ( tRoot ) );
		const auto tDone = std::chrono::steady_clock::now();
		m_tStats.host_wait_ms = fnMs ( tLaunched, tDone );
		m_tStats.d2h_bytes = (int64_t)( iOffCount + (size_t)nQueries*4 );
		for ( int s=0; s<nShards; ++s )
		{
			// (every shard's stream is done: the root's stream waited for all of them)
			cudaSetDevice ( m_dShards[s]->m_iDevice );
			dBatches[s]->Sync();
			const mgpu_batch_stats & t = dBatches[s]->m_tStats;
			m_tStats.h2d_bytes += t.h2d_bytes;
			m_tStats.kernel_launches += t.kernel_launches;
			m_tStats.max_eval_kernel_ms = std::max ( m_tStats.max_eval_kernel_ms, t.eval_kernel_ms );
			m_tStats.max_hot_decode_ms = std::max ( m_tStats.max_hot_decode_ms, t.hot_decode_ms );
			m_tStats.algorithmic_bytes += t.algorithmic_bytes;
			m_tStats.postings += t.postings;
		}
		m_tStats.kernel_launches += 1;

		const Key128_t * dKeys = (const Key128_t *)pStage;
		const int64_t * dDocid = (const int64_t *)( pStage+iOffDocid );
		const int64_t * dTotal = (const int64_t *)( pStage+iOffTotal );
		const int32_t * dCount = (const int32_t *)( pStage+iOffCount );
		auto fnUnpack = [&] ( int iFrom, int iTo )
		{
		for ( int i=iFrom; i<iTo; ++i )
		{
			const PlannedQuery_t & p = dTemplate[i];
			mgpu_result & r = pResults[i];
			r.status = p.m_iStatus;
			r.n_matches = 0;
			r.total_found = 0;
			if ( r.word_stats )
				for ( size_t w=0; w<p.m_dWordStats.size(); ++w )
					r.word_stats[w] = p.m_dWordStats[w];
			if ( p.m_iStatus!=MGPU_OK || !p.m_tDev.m_nOps )
				continue;
			r.n_matches = std::min ( dCount[i], p.m_tDev.m_iMaxMatches );
			for ( int s=0; s<nShards; ++s )
				r.total_found += dTotal[(size_t)s*nQueries+i];
			for ( int k=0; k<r.n_matches; ++k )
			{
				const Key128_t & tKey = dKeys[(size_t)i*iKMax+k];
				if ( r.rowid ) r.rowid[k] = ~(uint32_t)( tKey.m_uLo>>32 );	// global rowid = position in the unsharded index
				if ( r.weight ) r.weight[k] = (int32_t)(uint32_t)tKey.m_uLo;
				if ( r.docid ) r.docid[k] = dDocid[(size_t)i*iKMax+k];
				if ( r.sort_attr )
				{
					int64_t v = 0;
					if ( p.m_iFirstIntKeyShift>=0 )
					{
						uint64_t u = tKey.m_uHi>>p.m_iFirstIntKeyShift;
						if ( p.m_iFirstIntKeyBits==32 )
						{
							u &= 0xffffffffull;
							if ( !p.m_bFirstIntKeyDesc ) u = ~u & 0xffffffffull;
							v = (int64_t)u;
						} else
						{
							if ( !p.m_bFirstIntKeyDesc ) u = ~u;
							v = (int64_t)( u ^ 0x8000000000000000ull );
						}
					}
					r.sort_attr[k] = v;
				}
			}
		}
		};
		{
			const int nThreads = std::max ( 1, std::min ( { (int)std::thread::hardware_concurrency(), 8, nQueries/512 } ) );
			if ( nThreads<=1 )
				fnUnpack ( 0, nQueries );
			else
			{
				std::vector<std::thread> dThreads;
				for ( int t=0; t<nThreads; ++t )
					dThreads.emplace_back ( fnUnpack, (int)( (int64_t)nQueries*t/nThreads ), (int)( (int64_t)nQueries*( t+1 )/nThreads ) );
				for ( auto & t : dThreads )
					t.join();
			}
		}
		dBatches.clear();
		m_tStats.host_fetch_ms = fnMs ( tDone, std::chrono::steady_clock::now() );
		m_tStats.host_total_ms = fnMs ( tStart, std::chrono::steady_clock::now() );
		return MGPU_OK;
	}
};

} // namespace mgpu

using namespace mgpu;

struct mgpu_sharded { ShardedIndex_c m_t; };
static thread_local std::string g_sLastShardedOpenError;

extern "C"
{

int mgpu_sharded_open ( const char * const * path_prefixes, const int * devices, int n_shards, mgpu_sharded ** out )
{
	if ( !path_prefixes || !devices || n_shards<1 || n_shards>64 || !out )
		return MGPU_E_BAD_QUERY;
	*out = nullptr;
	for ( int s=0; s<n_shards; ++s )
		if ( !path_prefixes[s] )
			return MGPU_E_BAD_QUERY;
	std::unique_ptr<mgpu_sharded> p ( new mgpu_sharded );
	int iRes = p->m_t.Open ( path_prefixes, devices, n_shards );
	if ( iRes!=MGPU_OK )
	{
		g_sLastShardedOpenError = p->m_t.m_sError;
		return iRes;
	}
	*out = p.release();
	return MGPU_OK;
}

void mgpu_sharded_close ( mgpu_sharded * sh )
{
	delete sh;
}

int mgpu_sharded_search_batch ( mgpu_sharded * sh, const mgpu_query * queries, int n_queries, mgpu_result * results )
{
	if ( !sh || n_queries<0 || ( n_queries && ( !queries || !results ) ) )
		return MGPU_E_BAD_QUERY;
	std::lock_guard<std::mutex> tGuard ( sh->m_t.m_tLock );
	if ( !n_queries )
		return MGPU_OK;
	return sh->m_t.Search ( queries, n_queries, results );
}

int mgpu_sharded_set_option ( mgpu_sharded * sh, const char * name, int64_t value )
{
	if ( !sh || !name )
		return MGPU_E_BAD_QUERY;
	std::lock_guard<std::mutex> tGuard ( sh->m_t.m_tLock );
	for ( auto & p : sh->m_t.m_dShards )
		if ( !p->m_tOpt.Set ( name, value ) )
		{
			sh->m_t.m_sError = std::string ( "unknown option or value out of range: " ) + name;
			return MGPU_E_BAD_QUERY;
		}
	return MGPU_OK;
}

int mgpu_sharded_get_stats ( const mgpu_sharded * sh, mgpu_sharded_stats * out )
{
	if ( !sh || !out )
		return MGPU_E_BAD_QUERY;
	*out = sh->m_t.m_tStats;
	return MGPU_OK;
}

int64_t mgpu_sharded_total_docs ( const mgpu_sharded * sh )
{
	return sh ? sh->m_t.m_iTotalDocs : 0;
}

int mgpu_sharded_word_docs ( const mgpu_sharded * sh, const char * word, int64_t * docs )
{
	if ( !sh || !word )
		return 0;
	auto it = sh->m_t.m_hGlobalDocs.find ( sh->m_t.m_dShards[0]->DictKey ( word ) );
	if ( it==sh->m_t.m_hGlobalDocs.end() )
		return 0;
	if ( docs )
		*docs = it->second.m_iDocs;
	return 1;
}

int mgpu_sharded_word_stats ( const mgpu_sharded * sh, const char * word, int64_t * docs, int64_t * hits )
{
	if ( !sh || !word )
		return 0;
	auto it = sh->m_t.m_hGlobalDocs.find ( sh->m_t.m_dShards[0]->DictKey ( word ) );
	if ( it==sh->m_t.m_hGlobalDocs.end() )
		return 0;
	if ( docs )
		*docs = it->second.m_iDocs;
	if ( hits )
		*hits = it->second.m_iHits;
	return 1;
}

const char * mgpu_sharded_last_error ( const mgpu_sharded * sh )
{
	return sh ? sh->m_t.m_sError.c_str() : g_sLastShardedOpenError.c_str();
}

}
