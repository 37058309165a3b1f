// VByte ("zint"/"zoff") codec of index format v62 and little byte-buffer helpers.
// Reference: decoder src/fileio.cpp:31-45 (SPH_VARINT_DECODE), encoder src/sphinxstd.h:5545-5567
// (sphCalcZippedLen / sphZipValue): big-endian base-128, 0x80 = continuation, last byte MSB clear.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace mgpu
{

using BYTE = uint8_t;
using DWORD = uint32_t;
using RowID_t = uint32_t;
using Hitpos_t = uint32_t;
using SphOffset_t = int64_t;

static const RowID_t INVALID_ROWID = 0xFFFFFFFFu;
static const Hitpos_t EMPTY_HIT = 0;

// Hitman_c<8> (src/sphinx.h:768-827): field<<24 | end<<23 | pos
struct HITMAN
{
	static const DWORD FIELDEND_MASK = 1u<<23;
	static const DWORD POS_MASK = ( 1u<<23 )-1;
	static inline Hitpos_t Create ( int iField, int iPos )			{ return ( (DWORD)iField<<24 ) + ( (DWORD)iPos & POS_MASK ); }
	static inline int GetField ( Hitpos_t u )						{ return (int)( u>>24 ); }
	static inline int GetPos ( Hitpos_t u )							{ return (int)( u & POS_MASK ); }
	static inline bool IsEnd ( Hitpos_t u )							{ return ( u & FIELDEND_MASK )!=0; }
	static inline DWORD GetPosWithField ( Hitpos_t u )				{ return u & ~FIELDEND_MASK; }
};

inline int ZippedLen ( uint64_t v )
{
	int n = 1;
	v >>= 7;
	while ( v ) { v >>= 7; ++n; }
	return n;
}

/// growing byte buffer with the writer-side primitives of CSphWriter (src/fileio.cpp)
struct ByteBuf_t
{
	std::vector<BYTE> m_d;

	int64_t		Pos() const				{ return (int64_t)m_d.size(); }
	void		Truncate ( int64_t n )	{ m_d.resize ( (size_t)n ); }
	void		PutByte ( BYTE b )		{ m_d.push_back ( b ); }
	void		PutBytes ( const void * p, size_t n )	{ const BYTE * b = (const BYTE*)p; m_d.insert ( m_d.end(), b, b+n ); }
	void		PutDword ( DWORD v )	{ PutBytes ( &v, 4 ); }
	void		PutOffset ( int64_t v )	{ PutBytes ( &v, 8 ); }
	void		PutString ( const std::string & s )	{ PutDword ( (DWORD)s.size() ); if ( !s.empty() ) PutBytes ( s.data(), s.size() ); }
	void		Zip ( uint64_t v )
	{
		int n = ZippedLen ( v );
		for ( int i=n-1; i>=0; --i )
			m_d.push_back ( (BYTE)( ( 0x7f & ( v >> ( 7*i ) ) ) | ( i ? 0x80 : 0 ) ) );
	}
};

inline uint64_t UnzipAt ( const BYTE * & p )
{
	uint64_t v = 0;
	BYTE b;
	do { b = *p++; v = ( v<<7 ) + ( b & 0x7f ); } while ( b & 0x80 );
	return v;
}

/// bounded little-endian reader for .sph / .spi parsing
struct ByteReader_t
{
	const BYTE * m_p = nullptr;
	const BYTE * m_pEnd = nullptr;
	bool m_bError = false;

	ByteReader_t ( const BYTE * p, size_t n ) : m_p ( p ), m_pEnd ( p+n ) {}
	bool		Need ( size_t n )		{ if ( (size_t)( m_pEnd-m_p )<n ) { m_bError = true; return false; } return true; }
	BYTE		GetByte()				{ if ( !Need(1) ) return 0; return *m_p++; }
	DWORD		GetDword()				{ DWORD v = 0; if ( !Need(4) ) return 0; memcpy ( &v, m_p, 4 ); m_p += 4; return v; }
	int64_t		GetOffset()				{ int64_t v = 0; if ( !Need(8) ) return 0; memcpy ( &v, m_p, 8 ); m_p += 8; return v; }
	std::string	GetString()				{ DWORD n = GetDword(); if ( !Need(n) ) return std::string(); std::string s ( (const char*)m_p, n ); m_p += n; return s; }
	uint64_t	Unzip()					{ uint64_t v = 0; BYTE b; do { if ( !Need(1) ) return 0; b = *m_p++; v = ( v<<7 ) + ( b & 0x7f ); } while ( b & 0x80 ); return v; }
};

} // namespace mgpu
