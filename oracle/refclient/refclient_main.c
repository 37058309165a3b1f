/* Test harness around the REFERENCE's own C client (api/libsphinxclient/sphinxclient.c, SphinxAPI protocol 1.30), compiled from the
 * sources where they lie under /root/reference by oracle/Makefile into oracle/_ref/refclient (git-ignored, travels to the GPU box).
 * Test infrastructure, like the rest of oracle/: it connects to a loopback socket served by tests/test_api_wire.py, which hands every
 * packet to mgpu_api_handle, and prints what the reference client made of the replies.
 * usage: refclient PORT SCENARIO */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sphinxclient.h"

static void print_result ( sphinx_result * r )
{
	int i, a;
	if ( !r )
	{
		printf ( "NORESULT\n" );
		return;
	}
	printf ( "status %d\n", r->status );
	if ( r->error && *r->error )
		printf ( "error %s\n", r->error );
	if ( r->warning && *r->warning )
		printf ( "warning %s\n", r->warning );
	if ( r->status!=SEARCHD_OK && r->status!=SEARCHD_WARNING )
		return;
	for ( i=0; i<r->num_fields; i++ )
		printf ( "field %s\n", r->fields[i] );
	for ( a=0; a<r->num_attrs; a++ )
		printf ( "attr %s %d\n", r->attr_names[a], r->attr_types[a] );
	for ( i=0; i<r->num_matches; i++ )
	{
		printf ( "match %llu %d", (unsigned long long)sphinx_get_id ( r, i ), sphinx_get_weight ( r, i ) );
		for ( a=0; a<r->num_attrs; a++ )
			printf ( " %lld", (long long)sphinx_get_int ( r, i, a ) );
		printf ( "\n" );
	}
	printf ( "total %d found %d\n", r->total, r->total_found );
	for ( i=0; i<r->num_words; i++ )
		printf ( "word %s %d %d\n", r->words[i].word, r->words[i].docs, r->words[i].hits );
}

int main ( int argc, char ** argv )
{
	sphinx_client * c;
	const char * scenario;
	if ( argc<3 )
		return 2;
	scenario = argv[2];
	c = sphinx_create ( SPH_TRUE );
	sphinx_set_server ( c, "127.0.0.1", atoi ( argv[1] ) );
	sphinx_set_connect_timeout ( c, 5.0f );

	if ( !strcmp ( scenario, "default" ) )
	{
		print_result ( sphinx_query ( c, "hello world", "idx", NULL ) );
	} else if ( !strcmp ( scenario, "any_attr_desc" ) )
	{
		sphinx_set_match_mode ( c, SPH_MATCH_ANY );
		sphinx_set_sort_mode ( c, SPH_SORT_ATTR_DESC, "group_id" );
		sphinx_set_limits ( c, 1, 3, 50, 0 );
		print_result ( sphinx_query ( c, "hello there", "idx", "a comment" ) );
	} else if ( !strcmp ( scenario, "extended_sort_filter_weights" ) )
	{
		const sphinx_int64_t values[2] = { 3, 1 };
		const char * names[1] = { "title" };
		const int weights[1] = { 5 };
		sphinx_set_match_mode ( c, SPH_MATCH_EXTENDED2 );
		sphinx_set_sort_mode ( c, SPH_SORT_EXTENDED, "@weight DESC, group_id ASC" );
		sphinx_add_filter ( c, "group_id", 2, values, SPH_FALSE );
		sphinx_set_field_weights ( c, 1, names, weights );
		sphinx_set_ranking_mode ( c, SPH_RANK_BM25, NULL );
		print_result ( sphinx_query ( c, "hello | world | there", "idx", NULL ) );
	} else if ( !strcmp ( scenario, "phrase_range_idrange" ) )
	{
		sphinx_set_match_mode ( c, SPH_MATCH_PHRASE );
		sphinx_add_filter_range ( c, "stamp", 100, 160, SPH_TRUE );
		sphinx_set_id_range ( c, 3, 40 );
		sphinx_set_limits ( c, 0, 100, 100, 0 );
		print_result ( sphinx_query ( c, "hello world", "idx", NULL ) );
	} else if ( !strcmp ( scenario, "multi" ) )
	{
		sphinx_result * r;
		int i, n;
		sphinx_set_match_mode ( c, SPH_MATCH_EXTENDED2 );
		sphinx_set_ranking_mode ( c, SPH_RANK_WORDCOUNT, NULL );
		sphinx_add_query ( c, "\"hello world\"~3 | extra", "idx", NULL );
		sphinx_set_groupby ( c, "group_id", SPH_GROUPBY_ATTR, "@group desc" );
		sphinx_add_query ( c, "hello", "idx", NULL );
		sphinx_reset_groupby ( c );
		sphinx_add_query ( c, "hello | (world", "idx", NULL );
		sphinx_set_sort_mode ( c, SPH_SORT_ATTR_ASC, "stamp" );
		sphinx_add_query ( c, "@title hello -there", "idx", NULL );
		r = sphinx_run_queries ( c );
		n = sphinx_get_num_results ( c );
		if ( !r )
			printf ( "NORESULT %s\n", sphinx_error ( c ) );
		for ( i=0; r && i<n; i++ )
		{
			printf ( "query %d\n", i );
			print_result ( r+i );
		}
	} else if ( !strcmp ( scenario, "smoke" ) )
	{
		/* the calls of the client's own smoke test (api/libsphinxclient/test.c:58-75, 322-345, 396-416, 450-465) */
		const char * names[2] = { "title", "content" };
		const int weights[2] = { 100, 1 };
		const sphinx_int64_t group = 1;
		const char * queries[3] = { "is", "is test", "test number" };
		int i, n = 0;
		sphinx_keyword_info * k;
		sphinx_set_match_mode ( c, SPH_MATCH_EXTENDED2 );
		sphinx_set_sort_mode ( c, SPH_SORT_RELEVANCE, NULL );
		k = sphinx_build_keywords ( c, "hello test one", "test1", SPH_TRUE, &n );
		for ( i=0; k && i<n; i++ )
			printf ( "keyword %s %s %d %d\n", k[i].tokenized, k[i].normalized, k[i].num_docs, k[i].num_hits );
		for ( i=0; i<3; i++ )
		{
			sphinx_set_field_weights ( c, 2, names, weights );
			print_result ( sphinx_query ( c, queries[i], "test1", NULL ) );
		}
		sphinx_add_filter ( c, "group_id", 1, &group, SPH_FALSE );
		sphinx_set_field_weights ( c, 2, names, weights );
		print_result ( sphinx_query ( c, "is", "test1", NULL ) );
		sphinx_reset_filters ( c );
	} else if ( !strcmp ( scenario, "keywords_stats" ) )
	{
		int i, n = 0;
		sphinx_keyword_info * k = sphinx_build_keywords ( c, "hello world hello zzz", "idx", SPH_TRUE, &n );
		if ( !k )
			printf ( "NORESULT %s\n", sphinx_error ( c ) );
		for ( i=0; k && i<n; i++ )
			printf ( "keyword %s %s %d %d\n", k[i].tokenized, k[i].normalized, k[i].num_docs, k[i].num_hits );
	} else
		return 2;
	if ( sphinx_error ( c ) && *sphinx_error ( c ) )
		printf ( "client_error %s\n", sphinx_error ( c ) );
	sphinx_destroy ( c );
	return 0;
}
