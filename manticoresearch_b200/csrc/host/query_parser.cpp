// Query front-end: extended query syntax -> the flattened XQNode_t tree mgpu_search_batch takes.
//
// Restates XQParser_t (src/sphinxquery.cpp:1011-1830; grammar src/sphinxquery.y:57-160) and the tree fix-ups of XQParseHelper_c
// (src/sphinxquery.cpp:310-559) as a hand-written lexer + recursive-descent parser (the reference generates its parser with bison,
// which this image does not have), the query-mode tokenizer rules the lexer leans on (CSphTokenizerBase::CodepointArbitrationQ,
// src/sphinx.cpp:4655-4711) for the default charset_table + CJK unigrams, and the legacy match modes' rewrite into extended syntax
// (PrepareQueryEmulation, src/searchd.cpp:2141-2190).
// SENTENCE / PARAGRAPH are parsed (the evaluators refuse them).  Not restated: blended characters, multiform destinations, query token-filter plugins, zones, exact-form
// '=' (index_exact_words), wildcards.  Those return MGPU_E_UNSUPPORTED or parse as the plain text would.
#include "../../../include/mgpu.h"

#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <strings.h>
#include <ctype.h>
#include <string>
#include <vector>
#include <memory>
#include <unordered_set>
#include <algorithm>

namespace mgpu
{

enum Tok_e
{
	TOK_EOF = 0,
	TOK_KEYWORD = 256, TOK_NEAR, TOK_NOTNEAR, TOK_INT, TOK_FLOAT, TOK_FIELDLIMIT, TOK_BEFORE, TOK_MAYBE, TOK_SENTENCE, TOK_PARAGRAPH,
	TOK_ERROR = -1
};

/// XQLimitSpec_t (src/sphinxquery.h:66-130), fields 0..31 only
struct LimitSpec_t
{
	uint32_t	m_uFieldMask = 0xFFFFFFFFu;
	int			m_iFieldMaxPos = 0;
	bool		m_bFieldSpec = false;
};

/// XQKeyword_t; m_bNull = the reference's m_sWord.cstr()==NULL (stop word, overshort filler)
struct PWord_t
{
	std::string	m_sWord;
	bool		m_bNull = false;
	int			m_iAtomPos = 0;
	float		m_fBoost = 1.0f;
	bool		m_bFieldStart = false, m_bFieldEnd = false;
	bool		m_bExcluded = false;
};

/// XQNode_t
struct PNode_t
{
	int						m_iOp = MGPU_OP_AND;
	int						m_iOpArg = 0;
	bool					m_bPercentOp = false;
	bool					m_bNullOp = false;		// SPH_QUERY_NULL
	LimitSpec_t				m_tSpec;
	std::vector<PWord_t>	m_dWords;
	std::vector<PNode_t*>	m_dChildren;
	int						m_iDepth = 1;			// height of the subtree: the fix-ups and the planner recurse over it
};

struct Token_t
{
	int			m_iType = 0;
	PNode_t *	m_pNode = nullptr;
	int			m_iValue = 0;
	float		m_fValue = 0.f;
	int			m_iStrIndex = -1;
	uint32_t	m_uMask = 0;
	int			m_iMaxPos = 0;
};

static inline bool IsSpace ( int c )	{ return c==' ' || c=='\t' || c=='\r' || c=='\n'; }
static inline bool IsAlphaRef ( int c )	{ return ( c>='0' && c<='9' ) || ( c>='a' && c<='z' ) || ( c>='A' && c<='Z' ) || c=='-' || c=='_'; }	// sphIsAlpha
static inline bool IsModifier ( int c )	{ return c=='^' || c=='$'; }

//////////////////////////////////////////////////////////////////////////
// query-mode tokenizer
//////////////////////////////////////////////////////////////////////////

/// what the lexer needs of ISphTokenizer in query mode (Clone ( SPH_CLONE_QUERY_LIGHTWEIGHT ) + sphSetupQueryTokenizer):
/// default charset_table (0..9, A..Z->a..z, _, a..z, U+410..U+42F->U+430..U+44F, U+430..U+44F, U+401->U+451, U+451),
/// optional CJK unigrams (ngram_len=1), specials ()|-!@~"/^$<, backslash escapes, min_word_len with overshort counting
class QueryTokenizer_c
{
public:
	bool			m_bPhrase = false;
	int				m_iMinWordLen = 1;
	bool			m_bCjk = true;
	bool			m_bIndexMode = false;	///< the index tokenizer: no specials, no escapes

	void SetBuffer ( const char * p, int iLen )
	{
		m_pStart = m_pCur = p;
		m_pEnd = p+iLen;
		m_pTokStart = m_pTokEnd = nullptr;
		m_iOvershort = 0;
	}
	const char *	GetBufferPtr () const			{ return m_pCur; }
	const char *	GetBufferEnd () const			{ return m_pEnd; }
	void			SetBufferPtr ( const char * p )	{ m_pCur = p; }
	const char *	GetTokenStart () const			{ return m_pTokStart; }
	const char *	GetTokenEnd () const			{ return m_pTokEnd; }
	int				GetOvershortCount () const		{ return m_iOvershort; }
	bool			WasTokenSpecial () const		{ return m_bSpecial; }

	/// nullptr at the end of the buffer
	const char * GetToken ()
	{
		m_bSpecial = false;
		m_iOvershort = 0;
		m_sAccum.clear();
		int iAccumChars = 0;
		const char * pAccumStart = nullptr;
		while ( true )
		{
			const char * pCharStart = m_pCur;
			bool bEscaped = false;
			int iCode = -1;
			if ( m_pCur<m_pEnd )
			{
				if ( *m_pCur=='\\' && !m_bIndexMode )
				{
					// an escaped character never acts as a special (CodepointArbitrationQ, bWasEscaped)
					++m_pCur;
					bEscaped = true;
					if ( m_pCur>=m_pEnd )
						iCode = -1;
				}
				if ( m_pCur<m_pEnd )
					iCode = DecodeUtf8();
			}

			int iFolded = iCode<0 ? 0 : Fold ( iCode );
			const bool bSpecialChar = !m_bIndexMode && iCode>=0 && iCode<128 && strchr ( "()|-!@~\"/^$<", iCode )!=nullptr;
			bool bSpecial = bSpecialChar;
			if ( bSpecial )
			{
				const bool bDashInside = iAccumChars && iCode=='-';
				const BYTE_t uNext = m_pCur<m_pEnd ? (BYTE_t)*m_pCur : 0;
				if ( bEscaped || bDashInside
					|| ( iAccumChars && iCode=='$' && !IsBoundary ( uNext ) )
					|| ( m_bPhrase && iCode!='"' && !IsModifier ( iCode ) ) )
					bSpecial = false;	// a separator now
			}
			const bool bNgram = !bSpecialChar && iCode>=0 && m_bCjk && IsCjk ( iCode );

			if ( iCode<0 || ( !iFolded && !bNgram ) || bSpecial || bNgram )
			{
				// token boundary
				if ( iAccumChars )
				{
					if ( bSpecial || bNgram )
						m_pCur = pCharStart;	// comes back on the next call
					if ( iAccumChars<m_iMinWordLen )
					{
						++m_iOvershort;
						m_sAccum.clear();
						iAccumChars = 0;
						if ( iCode<0 )
							return nullptr;
						continue;
					}
					m_pTokStart = pAccumStart;
					m_pTokEnd = ( bSpecial || bNgram ) ? pCharStart : pCharStart;
					return m_sAccum.c_str();
				}
				if ( iCode<0 )
					return nullptr;
				if ( bSpecial )
				{
					m_sAccum.assign ( 1, (char)iCode );
					m_bSpecial = true;
					m_pTokStart = pCharStart;
					m_pTokEnd = m_pCur;
					return m_sAccum.c_str();
				}
				if ( bNgram )
				{
					m_sAccum.assign ( pCharStart+( bEscaped ? 1 : 0 ), m_pCur );
					m_pTokStart = pCharStart;
					m_pTokEnd = m_pCur;
					return m_sAccum.c_str();	// n-grams are not subject to min_word_len
				}
				continue;	// a separator in front of a token
			}

			if ( !iAccumChars )
				pAccumStart = pCharStart;
			AppendUtf8 ( iFolded );
			++iAccumChars;
		}
	}

private:
	typedef unsigned char BYTE_t;
	const char *	m_pStart = nullptr;
	const char *	m_pCur = nullptr;
	const char *	m_pEnd = nullptr;
	const char *	m_pTokStart = nullptr;
	const char *	m_pTokEnd = nullptr;
	int				m_iOvershort = 0;
	bool			m_bSpecial = false;
	std::string		m_sAccum;

	bool IsBoundary ( BYTE_t c ) const	// IsBoundary, src/sphinx.cpp:4564
	{
		return c==0 || IsSpace ( c ) || c=='"' || ( !m_bPhrase && ( c=='(' || c==')' || c=='|' ) );
	}
	static bool IsCjk ( int c )
	{
		return ( c>=0x2E80 && c<=0x9FFF ) || ( c>=0xAC00 && c<=0xD7AF ) || ( c>=0xF900 && c<=0xFAFF );
	}
	static int Fold ( int c )
	{
		if ( c>='0' && c<='9' ) return c;
		if ( c>='a' && c<='z' ) return c;
		if ( c>='A' && c<='Z' ) return c+32;
		if ( c=='_' ) return c;
		if ( c>=0x410 && c<=0x42F ) return c+0x20;
		if ( c>=0x430 && c<=0x44F ) return c;
		if ( c==0x401 || c==0x451 ) return 0x451;
		return 0;
	}
	int DecodeUtf8 ()
	{
		BYTE_t c = (BYTE_t)*m_pCur++;
		if ( c<0x80 )
			return c;
		int n = ( c>=0xF0 ) ? 3 : ( c>=0xE0 ) ? 2 : ( c>=0xC0 ) ? 1 : 0;
		int v = c & ( 0x3F>>n );
		if ( !n )
			return 0xFFFD;
		while ( n-- && m_pCur<m_pEnd && ( (BYTE_t)*m_pCur & 0xC0 )==0x80 )
			v = ( v<<6 ) | ( (BYTE_t)*m_pCur++ & 0x3F );
		return v;
	}
	void AppendUtf8 ( int c )
	{
		if ( c<0x80 )
			m_sAccum += (char)c;
		else if ( c<0x800 )
		{
			m_sAccum += (char)( 0xC0 | ( c>>6 ) );
			m_sAccum += (char)( 0x80 | ( c & 0x3F ) );
		} else if ( c<0x10000 )
		{
			m_sAccum += (char)( 0xE0 | ( c>>12 ) );
			m_sAccum += (char)( 0x80 | ( ( c>>6 ) & 0x3F ) );
			m_sAccum += (char)( 0x80 | ( c & 0x3F ) );
		} else
		{
			m_sAccum += (char)( 0xF0 | ( c>>18 ) );
			m_sAccum += (char)( 0x80 | ( ( c>>12 ) & 0x3F ) );
			m_sAccum += (char)( 0x80 | ( ( c>>6 ) & 0x3F ) );
			m_sAccum += (char)( 0x80 | ( c & 0x3F ) );
		}
	}
};

//////////////////////////////////////////////////////////////////////////
// parser
//////////////////////////////////////////////////////////////////////////

class Parser_c
{
public:
	std::string		m_sError, m_sWarning;
	int				m_iErrorCode = MGPU_OK;

	Parser_c ( const mgpu_parser_settings & s )
	{
		m_tTok.m_iMinWordLen = s.min_word_len>0 ? s.min_word_len : 1;
		m_tTok.m_bCjk = s.ngram_cjk!=0;
		for ( int i=0; i<s.n_fields; ++i )
			m_dFields.push_back ( s.field_names[i] ? s.field_names[i] : "" );
		for ( int i=0; i<s.n_stopwords; ++i )
			if ( s.stopwords[i] )
				m_hStopwords.insert ( s.stopwords[i] );
		m_iOvershortStep = s.overshort_step<0 ? 0 : ( s.overshort_step>1 ? 1 : s.overshort_step );
		m_bEmptyStopword = ( s.stopword_step==0 );
	}
	~Parser_c ()
	{
		for ( PNode_t * p : m_dAll )
			delete p;
	}

	/// XQParser_t::Parse, src/sphinxquery.cpp:1741-1830
	PNode_t * Parse ( const char * sQuery )
	{
		m_sQuery = sQuery ? sQuery : "";
		// the relaxed syntax option (:1752-1760): unknown fields are warnings, not errors
		m_bStopOnInvalid = true;
		if ( m_sQuery.compare ( 0, 9, "@@relaxed" )==0 && !IsAlphaRef ( m_sQuery.size()>9 ? m_sQuery[9] : 0 ) )
		{
			m_sQuery.erase ( 0, 9 );
			m_bStopOnInvalid = false;
		}
		m_dStateSpec.clear();
		m_dSpecPool.clear();
		m_dSpecPool.emplace_back ( new LimitSpec_t );
		m_dStateSpec.push_back ( m_dSpecPool.back().get() );
		m_tTok.SetBuffer ( m_sQuery.c_str(), (int)m_sQuery.size() );

		PNode_t * pRoot = nullptr;
		if ( Peek()!=TOK_EOF )
		{
			pRoot = ParseExpr();
			if ( m_iErrorCode==MGPU_OK && Peek()!=TOK_EOF )
				SyntaxError();
		}
		if ( m_iErrorCode!=MGPU_OK && !m_bEmpty )
			return nullptr;
		if ( m_iErrorCode!=MGPU_OK )
		{
			// nothing but separators: an empty query, not an error (m_bEmpty, :1801)
			m_iErrorCode = MGPU_OK;
			m_sError.clear();
			pRoot = nullptr;
		}
		return FixupTree ( pRoot );
	}

private:
	QueryTokenizer_c				m_tTok;
	std::string						m_sQuery;
	std::vector<std::string>		m_dFields;
	std::unordered_set<std::string>	m_hStopwords;
	std::vector<PNode_t*>			m_dAll;
	std::vector<LimitSpec_t*>		m_dStateSpec;
	std::vector<std::unique_ptr<LimitSpec_t>> m_dSpecPool;
	std::vector<std::string>		m_dIntTokens;
	std::vector<int>				m_dPhraseStar;	// positions of the `*` placeholders inside the phrase being lexed

	int		m_iAtomPos = 0;
	int		m_iPendingNulls = 0, m_iPendingType = 0;
	Token_t	m_tPendingToken;
	bool	m_bWasKeyword = false, m_bQuoted = false, m_bCheckNumber = false, m_bEmpty = true, m_bEmptyStopword = false;
	int		m_iQuorumQuote = -1, m_iQuorumFSlash = -1;
	int		m_iOvershortStep = 1;

	bool	m_bHaveTok = false;
	Token_t	m_tCur;
	static const int MAX_TREE_DEPTH = 512;
	int		m_iParenDepth = 0;
	bool	m_bStopOnInvalid = true;

	int Fail ( int iCode, const std::string & s )
	{
		if ( m_iErrorCode==MGPU_OK )
		{
			m_iErrorCode = iCode;
			m_sError = s;
		}
		return TOK_ERROR;
	}
	void SyntaxError ()
	{
		Fail ( MGPU_E_BAD_QUERY, "syntax error near '" + std::string ( m_tTok.GetBufferPtr() ? m_tTok.GetBufferPtr() : "" ) + "'" );
	}

	PNode_t * NewNode ( const LimitSpec_t & tSpec )
	{
		PNode_t * p = new PNode_t;
		p->m_tSpec = tSpec;
		m_dAll.push_back ( p );
		return p;
	}

	bool IsStopword ( const char * s ) const	{ return m_hStopwords.count ( s )!=0; }

	//////////////////////////////////////////////////////////////////////////
	// lexer

	/// XQParser_t::HandleModifiers, :1568-1598
	void HandleModifiers ( PWord_t & w )
	{
		const char * sTokStart = m_tTok.GetTokenStart();
		const char * sTokEnd = m_tTok.GetTokenEnd();
		if ( !sTokStart || !sTokEnd )
			return;
		const char * sQuery = m_sQuery.c_str();
		if ( sTokStart<sQuery || sTokStart>sQuery+m_sQuery.size() )
			return;		// the token came from the number buffer
		w.m_bFieldStart = ( sTokStart-sQuery )>0 && sTokStart[-1]=='^' && !( ( sTokStart-sQuery )>1 && sTokStart[-2]=='\\' );
		if ( sTokEnd[0]=='$' )
		{
			w.m_bFieldEnd = true;
			++sTokEnd;
		}
		if ( sTokEnd[0]=='^' && ( sTokEnd[1]=='.' || isdigit ( (unsigned char)sTokEnd[1] ) ) )
		{
			char * pEnd;
			float fBoost = (float)strtod ( sTokEnd+1, &pEnd );
			if ( ( sTokEnd+1 )!=pEnd )
			{
				w.m_fBoost = fBoost;
				m_tTok.SetBufferPtr ( pEnd );
			}
		}
	}

	/// XQParser_t::AddKeyword ( const char *, int ), :1601-1610
	PNode_t * AddKeyword ( const char * sKeyword )
	{
		PWord_t w;
		w.m_bNull = !sKeyword;
		if ( sKeyword )
			w.m_sWord = sKeyword;
		w.m_iAtomPos = m_iAtomPos;
		HandleModifiers ( w );
		PNode_t * pNode = NewNode ( *m_dStateSpec.back() );
		pNode->m_dWords.push_back ( w );
		return pNode;
	}

	/// XQParseHelper_c::AddField, :49-74
	bool AddField ( uint32_t & uMask, const char * sField, int iLen )
	{
		std::string sName ( sField, iLen );
		for ( size_t i=0; i<m_dFields.size(); ++i )
			if ( !strcasecmp ( m_dFields[i].c_str(), sName.c_str() ) )
			{
				if ( i>=32 )
				{
					Fail ( MGPU_E_UNSUPPORTED, "field limits beyond the first 32 fields are not supported" );
					return false;
				}
				uMask |= 1u<<i;
				return true;
			}
		if ( m_bStopOnInvalid )
		{
			Fail ( MGPU_E_BAD_QUERY, "no field '" + sName + "' found in schema" );
			return false;
		}
		m_sWarning = "no field '" + sName + "' found in schema";	// @@relaxed: the limit matches no field, its keywords are dropped below
		return true;
	}

	/// XQParseHelper_c::ParseFields, :77-214
	bool ParseFields ( uint32_t & uMask, int & iMaxPos, bool & bIgnore )
	{
		uMask = 0;
		iMaxPos = 0;
		bIgnore = false;
		const char * pPtr = m_tTok.GetBufferPtr();
		const char * pLast = m_tTok.GetBufferEnd();
		if ( pPtr==pLast )
			return true;

		bool bNegate = false, bBlock = false;
		if ( *pPtr=='!' )
		{
			bNegate = true;
			++pPtr;
			if ( *pPtr=='(' ) { bBlock = true; ++pPtr; }
		} else if ( *pPtr=='*' )
		{
			uMask = 0xFFFFFFFFu;
			m_tTok.SetBufferPtr ( pPtr+1 );
			return true;
		} else if ( *pPtr=='(' )
		{
			bBlock = true;
			++pPtr;
		}

		if ( !IsAlphaRef ( *pPtr ) )
		{
			bIgnore = true;
			m_tTok.SetBufferPtr ( pPtr );
			return true;
		}

		if ( !bBlock )
		{
			const char * pFieldStart = pPtr;
			while ( pPtr<pLast && IsAlphaRef ( *pPtr ) )
				++pPtr;
			if ( !AddField ( uMask, pFieldStart, (int)( pPtr-pFieldStart ) ) )
				return false;
			m_tTok.SetBufferPtr ( pPtr );
			if ( bNegate )
				uMask = ~uMask;
		} else
		{
			bool bOK = false;
			const char * pFieldStart = nullptr;
			while ( pPtr<pLast )
			{
				if ( IsAlphaRef ( *pPtr ) )
				{
					if ( !pFieldStart )
						pFieldStart = pPtr;
					++pPtr;
					continue;
				}
				if ( !pFieldStart )
				{
					Fail ( MGPU_E_BAD_QUERY, "error parsing field list: invalid field block operator syntax near '" + std::string ( pPtr, pLast ) + "'" );
					return false;
				} else if ( *pPtr==',' )
				{
					if ( !AddField ( uMask, pFieldStart, (int)( pPtr-pFieldStart ) ) )
						return false;
					pFieldStart = nullptr;
					++pPtr;
				} else if ( *pPtr==')' )
				{
					if ( !AddField ( uMask, pFieldStart, (int)( pPtr-pFieldStart ) ) )
						return false;
					m_tTok.SetBufferPtr ( ++pPtr );
					if ( bNegate )
						uMask = ~uMask;
					bOK = true;
					break;
				} else
				{
					Fail ( MGPU_E_BAD_QUERY, std::string ( "error parsing field list: invalid character '" ) + *pPtr + "' in field block operator" );
					return false;
				}
			}
			if ( !bOK )
			{
				Fail ( MGPU_E_BAD_QUERY, "error parsing field list: missing closing ')' in field block operator" );
				return false;
			}
		}

		if ( pPtr<pLast && pPtr[0]=='[' && isdigit ( (unsigned char)pPtr[1] ) )
		{
			const char * p = pPtr+1;
			while ( *p && isdigit ( (unsigned char)*p ) )
				++p;
			if ( *p!=']' )
				return true;
			iMaxPos = (int)strtoul ( pPtr+1, nullptr, 10 );
			m_tTok.SetBufferPtr ( p+1 );
		}
		return true;
	}

	static bool IsSpecialRef ( char c )	// XQParser_t::IsSpecial, :1011
	{
		return c=='(' || c==')' || c=='|' || c=='-' || c=='!' || c=='@' || c=='~' || c=='"' || c=='/';
	}

	/// XQParser_t::GetNumber, :1106-1178 (no blended characters or synonyms here)
	bool GetNumber ( const char * p )
	{
		int iDots = 0;
		const char * sToken = p;
		const char * sEnd = m_tTok.GetBufferEnd();
		while ( p<sEnd && ( isdigit ( (unsigned char)*p ) || *p=='.' ) )
		{
			iDots += ( *p=='.' );
			++p;
		}
		if ( iDots && ( iDots>1 || p-sToken==iDots ) )
			p = sToken;
		if ( iDots==1 && ( m_iQuorumQuote!=m_iQuorumFSlash || m_iQuorumQuote!=m_iAtomPos ) )
			p = sToken;

		const int NUMBER_BUF_LEN = 10;
		const char cNext = p<sEnd ? *p : '\0';
		if ( p>sToken && p-sToken<NUMBER_BUF_LEN
			&& !( cNext=='-' && !( p-sToken==1 && IsModifier ( p[-1] ) ) )
			&& ( cNext=='\0' || IsSpace ( cNext ) || IsSpecialRef ( cNext ) ) )
		{
			std::string sNumber ( sToken, p );
			m_tPendingToken = Token_t();
			if ( iDots )
				m_tPendingToken.m_fValue = (float)strtod ( sNumber.c_str(), nullptr );
			else
				m_tPendingToken.m_iValue = atoi ( sNumber.c_str() );

			// can it be a keyword too?
			QueryTokenizer_c tNum;
			tNum.m_iMinWordLen = m_tTok.m_iMinWordLen;
			tNum.m_bCjk = m_tTok.m_bCjk;
			tNum.SetBuffer ( sNumber.c_str(), (int)sNumber.size() );
			const char * sKw = tNum.GetToken();
			m_tTok.SetBufferPtr ( p );
			m_tPendingToken.m_iStrIndex = -1;
			if ( sKw )
			{
				if ( !IsStopword ( sKw ) )
				{
					m_dIntTokens.push_back ( sKw );
					m_tPendingToken.m_iStrIndex = (int)m_dIntTokens.size()-1;
				}
				m_iAtomPos++;
			}
			m_iPendingNulls = 0;
			m_iPendingType = iDots ? TOK_FLOAT : TOK_INT;
			return true;
		}
		return false;
	}

	bool GetNearToken ( const char * sTok, int iTokLen, int iTokType, const char * sBuf )
	{
		const char * sEnd = m_tTok.GetBufferEnd();
		if ( sEnd-sBuf>iTokLen && strncmp ( sBuf, sTok, iTokLen )==0 && isdigit ( (unsigned char)sBuf[iTokLen] ) )
		{
			int iVal = 0;
			for ( sBuf += iTokLen; sBuf<sEnd && isdigit ( (unsigned char)*sBuf ); ++sBuf )
				iVal = iVal*10 + ( *sBuf )-'0';
			m_tTok.SetBufferPtr ( sBuf );
			m_iPendingType = iTokType;
			m_tPendingToken = Token_t();
			m_tPendingToken.m_iValue = iVal;
			return true;
		}
		return false;
	}

	/// XQParser_t::GetToken, :1201-1554
	int GetToken ( Token_t & tOut )
	{
		if ( !m_iPendingType )
			while ( true )
		{
			const bool bWasKeyword = m_bWasKeyword;
			m_bWasKeyword = false;

			const char * pTokenStart = m_tTok.GetBufferPtr();
			const char * sBufferEnd = m_tTok.GetBufferEnd();
			const char * p = pTokenStart;
			while ( p<sBufferEnd && isspace ( (unsigned char)*p ) )
				++p;

			if ( m_bCheckNumber )
			{
				m_bCheckNumber = false;
				if ( GetNumber ( p ) )
					break;
			}

			const char * pLastTokenEnd = m_tTok.GetTokenEnd();
			const char * sToken = m_tTok.GetToken();
			if ( !sToken )
			{
				m_iPendingNulls = m_tTok.GetOvershortCount()*m_iOvershortStep;
				if ( !( m_iPendingNulls || m_tTok.GetBufferPtr()-p>0 ) )
					return TOK_EOF;
				m_iPendingNulls = 0;
				tOut = Token_t();
				tOut.m_pNode = AddKeyword ( nullptr );
				m_bWasKeyword = true;
				return TOK_KEYWORD;
			}
			m_bEmpty = false;

			m_iPendingNulls = m_tTok.GetOvershortCount()*m_iOvershortStep;
			m_iAtomPos += 1+m_iPendingNulls;

			const bool bPhrase = m_tTok.m_bPhrase;
			if ( !bPhrase && ( GetNearToken ( "NEAR/", 5, TOK_NEAR, p ) || GetNearToken ( "NOTNEAR/", 8, TOK_NOTNEAR, p ) ) )
			{
				m_iAtomPos -= 1;
				break;
			}
			const size_t nLeft = (size_t)( sBufferEnd-p );
			if ( !bPhrase && !strcasecmp ( sToken, "sentence" ) && nLeft>=8 && !strncmp ( p, "SENTENCE", 8 ) )
			{
				m_iPendingType = TOK_SENTENCE;
				m_tPendingToken = Token_t();
				m_iAtomPos -= 1;
				break;
			}
			if ( !bPhrase && !strcasecmp ( sToken, "paragraph" ) && nLeft>=9 && !strncmp ( p, "PARAGRAPH", 9 ) )
			{
				m_iPendingType = TOK_PARAGRAPH;
				m_tPendingToken = Token_t();
				m_iAtomPos -= 1;
				break;
			}
			if ( !bPhrase && !strcasecmp ( sToken, "maybe" ) && nLeft>=5 && !strncmp ( p, "MAYBE", 5 ) )
			{
				m_iPendingType = TOK_MAYBE;
				m_tPendingToken = Token_t();
				m_iAtomPos -= 1;
				break;
			}
			if ( !bPhrase && ( ( nLeft>5 && !strncmp ( p, "ZONE:", 5 ) && ( IsAlphaRef ( p[5] ) || p[5]=='(' ) )
				|| ( nLeft>9 && !strncmp ( p, "ZONESPAN:", 9 ) && ( IsAlphaRef ( p[9] ) || p[9]=='(' ) ) ) )
				return Fail ( MGPU_E_UNSUPPORTED, "ZONE / ZONESPAN limits are not supported" );

			// a separate star inside a phrase ("that * box") shifts the in-query positions behind it: count the [ * ] between the tokens (:1318-1348)
			if ( bPhrase && pLastTokenEnd && m_tTok.GetTokenStart() )
			{
				int iSpace = 0, iStar = 0;
				for ( const char * sCur = pLastTokenEnd; sCur<m_tTok.GetTokenStart(); ++sCur )
				{
					const int iCur = (int)( sCur-pLastTokenEnd );
					if ( *sCur=='*' )
						iStar = iCur;
					else if ( *sCur==' ' )
					{
						if ( iSpace+2==iCur && iStar+1==iCur )
							m_dPhraseStar.push_back ( m_iAtomPos );
						iSpace = iCur;
					}
				}
			}

			if ( m_tTok.WasTokenSpecial() )
			{
				m_iAtomPos--;	// specials must not affect pos
				if ( sToken[0]=='@' )
				{
					bool bIgnore;
					m_tPendingToken = Token_t();
					if ( !ParseFields ( m_tPendingToken.m_uMask, m_tPendingToken.m_iMaxPos, bIgnore ) )
						return TOK_ERROR;
					if ( bIgnore )
						continue;
					m_iPendingType = TOK_FIELDLIMIT;
					break;
				} else if ( sToken[0]=='<' )
				{
					if ( m_tTok.GetBufferPtr()<sBufferEnd && *m_tTok.GetBufferPtr()=='<' )
					{
						m_iPendingType = TOK_BEFORE;
						m_tPendingToken = Token_t();
						break;
					}
					if ( m_iPendingNulls>0 )
					{
						m_iPendingNulls = 0;
						tOut = Token_t();
						tOut.m_pNode = AddKeyword ( nullptr );
						m_bWasKeyword = true;
						return TOK_KEYWORD;
					}
					continue;
				} else if ( sToken[0]=='^' )
				{
					continue;	// HandleModifiers' business
				} else if ( sToken[0]=='$' )
				{
					if ( bWasKeyword )
						continue;
					if ( m_tTok.GetTokenStart()>m_sQuery.c_str() && IsSpace ( m_tTok.GetTokenStart()[-1] ) )
						continue;
					if ( m_tTok.GetOvershortCount()==1 )
					{
						m_iPendingNulls = 0;
						tOut = Token_t();
						tOut.m_pNode = AddKeyword ( nullptr );
						return TOK_KEYWORD;
					}
					m_sWarning = "modifiers must be applied to keywords, not operators";
					continue;
				} else
				{
					const bool bWasQuoted = m_bQuoted;
					if ( sToken[0]=='"' )
					{
						m_bQuoted = !m_bQuoted;
						if ( m_bQuoted )
							m_dPhraseStar.clear();
					}
					m_iPendingType = sToken[0];
					m_tPendingToken = Token_t();
					m_tTok.m_bPhrase = m_bQuoted;

					if ( sToken[0]=='(' )
						m_dStateSpec.push_back ( m_dStateSpec.back() );
					else if ( sToken[0]==')' && m_dStateSpec.size()>1 )
						m_dStateSpec.pop_back();

					if ( bWasQuoted && !m_bQuoted )
						m_iQuorumQuote = m_iAtomPos;
					else if ( sToken[0]=='/' )
						m_iQuorumFSlash = m_iAtomPos;

					if ( sToken[0]=='~' || sToken[0]=='/' )
						m_bCheckNumber = true;
					break;
				}
			}

			// a stop word keeps its position but has no keyword (GetWordID returns 0)
			std::string sWord ( sToken );
			const char * sKw = sWord.c_str();
			if ( IsStopword ( sKw ) )
			{
				sKw = nullptr;
				if ( m_bEmptyStopword )
					m_iAtomPos--;
			}
			m_tPendingToken = Token_t();
			m_tPendingToken.m_pNode = AddKeyword ( sKw );
			m_iPendingType = TOK_KEYWORD;
			break;
		}

		m_bEmpty = false;
		if ( m_iPendingNulls>0 )
		{
			m_iPendingNulls--;
			tOut = Token_t();
			tOut.m_pNode = AddNullKeyword();
			m_bWasKeyword = true;
			return TOK_KEYWORD;
		}

		int iRes = m_iPendingType;
		m_iPendingType = 0;
		if ( iRes==TOK_KEYWORD )
			m_bWasKeyword = true;
		tOut = m_tPendingToken;
		tOut.m_iType = iRes;
		return iRes;
	}

	/// the fillers for overshort tokens are plain nulls: no modifiers (the token pointers belong to the real token behind them)
	PNode_t * AddNullKeyword ()
	{
		PWord_t w;
		w.m_bNull = true;
		w.m_iAtomPos = m_iAtomPos;
		PNode_t * pNode = NewNode ( *m_dStateSpec.back() );
		pNode->m_dWords.push_back ( w );
		return pNode;
	}

	// lazy one-token lookahead: bison takes its default reductions without reading a token, so a token is lexed only when a decision needs it
	int Peek ()
	{
		if ( m_iErrorCode!=MGPU_OK )
			return TOK_ERROR;
		if ( !m_bHaveTok )
		{
			m_tCur = Token_t();
			m_tCur.m_iType = GetToken ( m_tCur );
			if ( m_iErrorCode!=MGPU_OK )
				m_tCur.m_iType = TOK_ERROR;
			m_bHaveTok = true;
		}
		return m_tCur.m_iType;
	}
	Token_t Take ()
	{
		Peek();
		m_bHaveTok = false;
		return m_tCur;
	}

	//////////////////////////////////////////////////////////////////////////
	// grammar, src/sphinxquery.y:57-160

	/// XQParser_t::SetFieldSpec + FixRefSpec
	void SetFieldSpec ( uint32_t uMask, int iMaxPos )
	{
		const size_t n = m_dStateSpec.size();
		if ( n>1 && m_dStateSpec[n-1]==m_dStateSpec[n-2] )
		{
			m_dSpecPool.emplace_back ( new LimitSpec_t ( *m_dStateSpec.back() ) );
			m_dStateSpec.back() = m_dSpecPool.back().get();
		}
		LimitSpec_t * p = m_dStateSpec.back();
		p->m_bFieldSpec = true;
		p->m_uFieldMask = uMask;
		p->m_iFieldMaxPos = iMaxPos;
	}

	static bool HasMissedField ( const LimitSpec_t & t )	{ return t.m_uFieldMask==0 && t.m_iFieldMaxPos==0; }

	/// XQParser_t::AddOp, :1634-1678
	PNode_t * AddOp ( int iOp, PNode_t * pLeft, PNode_t * pRight, int iOpArg=0 )
	{
		if ( iOp==MGPU_OP_NOT )
		{
			PNode_t * pNode = NewNode ( *m_dStateSpec.back() );
			pNode->m_iOp = MGPU_OP_NOT;
			if ( pLeft )
			{
				pNode->m_dChildren.push_back ( pLeft );
				pNode->m_iDepth = pLeft->m_iDepth+1;
			}
			return pNode;
		}
		if ( !pLeft || !pRight )
			return pLeft ? pLeft : pRight;
		if ( !pLeft->m_dChildren.empty() && pLeft->m_iOp==iOp && pLeft->m_iOpArg==iOpArg )
		{
			pLeft->m_dChildren.push_back ( pRight );
			pLeft->m_iDepth = std::max ( pLeft->m_iDepth, pRight->m_iDepth+1 );
			return pLeft;
		}
		PNode_t * pNode = NewNode ( HasMissedField ( pRight->m_tSpec ) ? pLeft->m_tSpec : pRight->m_tSpec );
		pNode->m_iOp = iOp;
		pNode->m_iOpArg = iOpArg;
		pNode->m_dChildren.push_back ( pLeft );
		pNode->m_dChildren.push_back ( pRight );
		pNode->m_iDepth = std::max ( pLeft->m_iDepth, pRight->m_iDepth )+1;
		if ( pNode->m_iDepth>MAX_TREE_DEPTH )
			Fail ( MGPU_E_BAD_QUERY, "query too complex, not enough stack (tree deeper than " + std::to_string ( MAX_TREE_DEPTH ) + " levels)" );
		return pNode;
	}

	bool StartsLimiterOrAtom ( int t ) const
	{
		return t==TOK_KEYWORD || t==TOK_INT || t==TOK_FLOAT || t==TOK_FIELDLIMIT || t=='"' || t=='(';
	}

	void TokLimiter ()
	{
		if ( Peek()==TOK_FIELDLIMIT )
		{
			Token_t t = Take();
			SetFieldSpec ( t.m_uMask, t.m_iMaxPos );
		}
	}

	/// expr: beforelist | expr beforelist
	PNode_t * ParseExpr ()
	{
		PNode_t * pRes = ParseBeforeList();
		while ( m_iErrorCode==MGPU_OK )
		{
			int t = Peek();
			if ( !( StartsLimiterOrAtom ( t ) || t=='-' || t=='!' ) )
				break;
			PNode_t * pNext = ParseBeforeList();
			pRes = AddOp ( MGPU_OP_AND, pRes, pNext );
		}
		return pRes;
	}

	/// beforelist: orlistf | beforelist TOK_BEFORE orlistf | beforelist TOK_NEAR orlistf
	PNode_t * ParseBeforeList ()
	{
		PNode_t * pRes = ParseOrListF();
		while ( m_iErrorCode==MGPU_OK )
		{
			int t = Peek();
			if ( t==TOK_BEFORE )
			{
				Take();
				pRes = AddOp ( MGPU_OP_BEFORE, pRes, ParseOrListF() );
			} else if ( t==TOK_NEAR )
			{
				Token_t tNear = Take();
				pRes = AddOp ( MGPU_OP_NEAR, pRes, ParseOrListF(), tNear.m_iValue );
			} else
				break;
		}
		return pRes;
	}

	/// orlistf: orlist | tok_limiter '-' orlist | tok_limiter '!' orlist
	PNode_t * ParseOrListF ()
	{
		TokLimiter();
		int t = Peek();
		if ( t=='-' || t=='!' )
		{
			Take();
			return AddOp ( MGPU_OP_NOT, ParseOrList ( false ), nullptr );
		}
		return ParseOrList ( true );
	}

	/// orlist: tok_limiter atom | orlist '|' tok_limiter atom | orlist TOK_MAYBE tok_limiter atom
	PNode_t * ParseOrList ( bool bLimiterDone )
	{
		if ( !bLimiterDone )
			TokLimiter();
		bool bOk = false;
		PNode_t * pRes = ParseAtom ( bOk );
		if ( !bOk )
			return nullptr;
		while ( m_iErrorCode==MGPU_OK )
		{
			int t = Peek();
			if ( t!='|' && t!=TOK_MAYBE )
				break;
			Take();
			TokLimiter();
			PNode_t * pNext = ParseAtom ( bOk );
			if ( !bOk )
				return nullptr;
			pRes = AddOp ( t=='|' ? MGPU_OP_OR : MGPU_OP_MAYBE, pRes, pNext );
		}
		return pRes;
	}

	/// atom (with the left-associative `atom TOK_NOTNEAR atom`)
	PNode_t * ParseAtom ( bool & bOk )
	{
		PNode_t * pRes = ParsePrimary ( bOk );
		while ( bOk && m_iErrorCode==MGPU_OK && Peek()==TOK_NOTNEAR )
		{
			Token_t t = Take();
			PNode_t * pRight = ParsePrimary ( bOk );
			if ( !bOk )
				return nullptr;
			pRes = AddOp ( MGPU_OP_NOTNEAR, pRes, pRight, t.m_iValue );
		}
		return pRes;
	}

	/// XQParser_t::PhraseShiftQpos, :1701-1738
	void PhraseShiftQpos ( PNode_t * pNode )
	{
		if ( m_dPhraseStar.empty() )
			return;
		size_t iLast = 0;
		const int iQposShiftStart = m_dPhraseStar[0];
		int iQposShift = 0;
		int iLastStarPos = m_dPhraseStar[0];
		for ( size_t iWord=0; iWord<pNode->m_dWords.size(); ++iWord )
		{
			PWord_t & tWord = pNode->m_dWords[iWord];
			while ( iLast<m_dPhraseStar.size() && m_dPhraseStar[iLast]<=tWord.m_iAtomPos )
			{
				iLastStarPos = m_dPhraseStar[iLast];
				++iLast;
				++iQposShift;
			}
			if ( tWord.m_sWord=="*" || ( tWord.m_bNull && tWord.m_iAtomPos==iLastStarPos ) )
			{
				pNode->m_dWords.erase ( pNode->m_dWords.begin()+iWord );
				--iWord;
				--iQposShift;
				continue;
			}
			if ( iQposShiftStart<=tWord.m_iAtomPos )
				tWord.m_iAtomPos += iQposShift;
		}
	}

	/// sentence: sp_item TOK_SENTENCE sp_item | sentence TOK_SENTENCE sp_item (and the same for paragraph); sp_item: keyword | '"' phrase '"'
	PNode_t * UnitChain ( PNode_t * pLeft, bool & bOk )
	{
		const int iFirst = Peek();
		if ( iFirst!=TOK_SENTENCE && iFirst!=TOK_PARAGRAPH )
			return pLeft;
		while ( m_iErrorCode==MGPU_OK && Peek()==iFirst )
		{
			Take();
			PNode_t * pRight = nullptr;
			const int t = Peek();
			if ( t==TOK_KEYWORD )
				pRight = Take().m_pNode;
			else if ( t==TOK_INT || t==TOK_FLOAT )
				pRight = KeywordOfNumber ( Take() );
			else if ( t=='"' )
			{
				Take();
				while ( m_iErrorCode==MGPU_OK )
				{
					const int k = Peek();
					PNode_t * pTok = nullptr;
					if ( k==TOK_KEYWORD )
						pTok = Take().m_pNode;
					else if ( k==TOK_INT || k==TOK_FLOAT )
						pTok = KeywordOfNumber ( Take() );
					else
						break;
					if ( pTok )
					{
						if ( !pRight )
							pRight = pTok;
						else
							pRight->m_dWords.push_back ( pTok->m_dWords[0] );
					}
				}
				if ( Peek()!='"' )
				{
					bOk = false;
					SyntaxError();
					return nullptr;
				}
				Take();
				if ( pRight )
					pRight->m_iOp = MGPU_OP_PHRASE;
			} else
			{
				bOk = false;
				SyntaxError();
				return nullptr;
			}
			pLeft = AddOp ( iFirst==TOK_SENTENCE ? MGPU_OP_SENTENCE : MGPU_OP_PARAGRAPH, pLeft, pRight );
		}
		const int iNext = Peek();
		if ( iNext==TOK_SENTENCE || iNext==TOK_PARAGRAPH )
		{
			bOk = false;	// the grammar does not mix the two in one chain
			SyntaxError();
			return nullptr;
		}
		return pLeft;
	}

	PNode_t * KeywordOfNumber ( const Token_t & t )
	{
		return AddKeyword ( t.m_iStrIndex>=0 ? m_dIntTokens[t.m_iStrIndex].c_str() : nullptr );
	}

	PNode_t * ParsePrimary ( bool & bOk )
	{
		bOk = true;
		int t = Peek();
		if ( t==TOK_KEYWORD )
			return UnitChain ( Take().m_pNode, bOk );
		if ( t==TOK_INT || t==TOK_FLOAT )
			return UnitChain ( KeywordOfNumber ( Take() ), bOk );
		if ( t=='(' )
		{
			Take();
			if ( ++m_iParenDepth>MAX_TREE_DEPTH )
			{
				// the reference measures its stack (sphGetStackUsed) and answers "query too complex"; here the nesting is counted
				bOk = false;
				Fail ( MGPU_E_BAD_QUERY, "query too complex, not enough stack (parentheses nested deeper than " + std::to_string ( MAX_TREE_DEPTH ) + " levels)" );
				return nullptr;
			}
			PNode_t * pRes = ParseExpr();
			--m_iParenDepth;
			if ( Peek()!=')' )
			{
				bOk = false;
				SyntaxError();
				return nullptr;
			}
			Take();
			return pRes;
		}
		if ( t=='"' )
		{
			Take();
			// phrase: phrasetoken+ ; specials inside the quotes come as separators from the tokenizer
			PNode_t * pPhrase = nullptr;
			bool bAny = false;
			while ( m_iErrorCode==MGPU_OK )
			{
				int k = Peek();
				PNode_t * pTok = nullptr;
				if ( k==TOK_KEYWORD )
					pTok = Take().m_pNode;
				else if ( k==TOK_INT || k==TOK_FLOAT )
					pTok = KeywordOfNumber ( Take() );
				else if ( k=='(' || k==')' || k=='-' || k=='|' || k=='~' || k=='/' )
					Take();
				else
					break;
				bAny = true;
				if ( pTok )
				{
					if ( !pPhrase )
						pPhrase = pTok;
					else
						pPhrase->m_dWords.push_back ( pTok->m_dWords[0] );	// AddKeyword ( left, right ), :1613
				}
			}
			(void)bAny;
			if ( Peek()!='"' )
			{
				bOk = false;
				SyntaxError();
				return nullptr;
			}
			Take();
			int k = Peek();
			if ( k=='~' || k=='/' )
			{
				Take();
				int n = Peek();
				if ( !( n==TOK_INT || ( k=='/' && n==TOK_FLOAT ) ) )
				{
					bOk = false;
					SyntaxError();
					return nullptr;
				}
				Token_t tNum = Take();
				if ( pPhrase )
				{
					if ( k=='~' )
					{
						pPhrase->m_iOp = MGPU_OP_PROXIMITY;
						pPhrase->m_iOpArg = tNum.m_iValue;
						m_iAtomPos = pPhrase->m_dWords.back().m_iAtomPos+1;	// XQNode_t::FixupAtomPos (nothing skipped: no blended parts)
					} else if ( n==TOK_INT )
					{
						pPhrase->m_iOp = MGPU_OP_QUORUM;
						pPhrase->m_iOpArg = tNum.m_iValue;
					} else
					{
						pPhrase->m_iOp = MGPU_OP_QUORUM;
						pPhrase->m_iOpArg = (int)( tNum.m_fValue*100 );
						pPhrase->m_bPercentOp = true;
					}
				}
				return pPhrase;
			}
			const int iUnit = Peek();
			if ( iUnit==TOK_SENTENCE || iUnit==TOK_PARAGRAPH )
			{
				if ( pPhrase )
					pPhrase->m_iOp = MGPU_OP_PHRASE;	// sp_item: '"' phrase '"' (no star shift there)
				return UnitChain ( pPhrase, bOk );
			}
			if ( pPhrase )
			{
				pPhrase->m_iOp = MGPU_OP_PHRASE;	// SetPhrase
				PhraseShiftQpos ( pPhrase );
			}
			return pPhrase;
		}
		bOk = false;
		SyntaxError();
		return nullptr;
	}

	//////////////////////////////////////////////////////////////////////////
	// fix-ups, src/sphinxquery.cpp:310-559

	PNode_t * SweepNulls ( PNode_t * pNode )
	{
		if ( !pNode )
			return nullptr;
		if ( !pNode->m_dWords.empty() )
		{
			std::vector<PWord_t> dKeep;
			for ( const auto & w : pNode->m_dWords )
				if ( !w.m_bNull )
					dKeep.push_back ( w );
			pNode->m_dWords.swap ( dKeep );
			return pNode->m_dWords.empty() ? nullptr : pNode;
		}
		for ( size_t i=0; i<pNode->m_dChildren.size(); )
		{
			pNode->m_dChildren[i] = SweepNulls ( pNode->m_dChildren[i] );
			if ( !pNode->m_dChildren[i] )
			{
				pNode->m_dChildren.erase ( pNode->m_dChildren.begin()+i );
				++pNode->m_iOpArg;		// the reference's "sweeping happened" flag
			} else
				++i;
		}
		if ( pNode->m_dChildren.empty() )
			return nullptr;
		if ( pNode->m_iOp!=MGPU_OP_NOT && pNode->m_dChildren.size()==1 )
		{
			PNode_t * pRet = pNode->m_dChildren[0];
			pNode->m_dChildren.clear();
			if ( pNode->m_iOpArg && pRet->m_iOp==MGPU_OP_NOT && !pRet->m_bNullOp )
			{
				pRet->m_bNullOp = true;
				pRet->m_dChildren.clear();
			}
			pRet->m_iOpArg = pNode->m_iOpArg;
			return SweepNulls ( pRet );
		}
		return pNode;
	}

	void FixupDegenerates ( PNode_t * pNode )
	{
		if ( !pNode )
			return;
		if ( pNode->m_dWords.size()==1 && ( pNode->m_iOp==MGPU_OP_PHRASE || pNode->m_iOp==MGPU_OP_PROXIMITY || pNode->m_iOp==MGPU_OP_QUORUM ) )
		{
			if ( pNode->m_iOp==MGPU_OP_QUORUM && !pNode->m_bPercentOp && pNode->m_iOpArg>1 )
				m_sWarning = "quorum threshold too high (words=1, thresh=" + std::to_string ( pNode->m_iOpArg ) + "); replacing quorum operator with AND operator";
			pNode->m_iOp = MGPU_OP_AND;
			return;
		}
		for ( PNode_t * p : pNode->m_dChildren )
			FixupDegenerates ( p );
	}

	void FixupNulls ( PNode_t * pNode )
	{
		if ( !pNode )
			return;
		for ( PNode_t * p : pNode->m_dChildren )
			FixupNulls ( p );
		if ( pNode->m_bNullOp )
			return;
		if ( pNode->m_iOp==MGPU_OP_OR )
		{
			std::vector<PNode_t*> dKeep;
			for ( PNode_t * p : pNode->m_dChildren )
				if ( !p->m_bNullOp )
					dKeep.push_back ( p );
			pNode->m_dChildren.swap ( dKeep );
		} else if ( pNode->m_iOp==MGPU_OP_AND && pNode->m_dWords.empty() )
		{
			for ( PNode_t * p : pNode->m_dChildren )
				if ( p->m_bNullOp )
				{
					pNode->m_bNullOp = true;
					pNode->m_dChildren.clear();
					break;
				}
		}
	}

	bool FixupNots ( PNode_t * pNode )
	{
		if ( !pNode || !pNode->m_dWords.empty() )
			return true;
		for ( PNode_t * p : pNode->m_dChildren )
			if ( !FixupNots ( p ) )
				return false;

		std::vector<PNode_t*> dNots;
		for ( size_t i=0; i<pNode->m_dChildren.size(); ++i )
			if ( pNode->m_dChildren[i]->m_iOp==MGPU_OP_NOT && !pNode->m_dChildren[i]->m_bNullOp )
			{
				dNots.push_back ( pNode->m_dChildren[i] );
				pNode->m_dChildren[i] = pNode->m_dChildren.back();	// RemoveFast
				pNode->m_dChildren.pop_back();
				--i;
			}
		if ( dNots.empty() )
			return true;
		if ( pNode->m_dChildren.empty() )
		{
			Fail ( MGPU_E_BAD_QUERY, "query is non-computable (node consists of NOT operators only)" );
			return false;
		}
		if ( pNode->m_iOp==MGPU_OP_OR || pNode->m_iOp==MGPU_OP_MAYBE || pNode->m_iOp==MGPU_OP_NEAR )
		{
			const char * sOp = pNode->m_iOp==MGPU_OP_OR ? "OR" : ( pNode->m_iOp==MGPU_OP_MAYBE ? "MAYBE" : "NEAR" );
			Fail ( MGPU_E_BAD_QUERY, std::string ( "query is non-computable (NOT is not allowed within " ) + sOp + ")" );
			return false;
		}
		if ( pNode->m_iOp==MGPU_OP_BEFORE )
		{
			Fail ( MGPU_E_BAD_QUERY, "query is non-computable (NOT cannot be used as before operand)" );
			return false;
		}
		if ( pNode->m_iOp!=MGPU_OP_AND )
		{
			Fail ( MGPU_E_BAD_QUERY, "query is non-computable (NOT inside this operator)" );
			return false;
		}

		PNode_t * pAnd = NewNode ( pNode->m_tSpec );
		pAnd->m_iOp = MGPU_OP_AND;
		pAnd->m_dChildren = pNode->m_dChildren;
		PNode_t * pNot;
		if ( dNots.size()==1 )
			pNot = dNots[0];
		else
		{
			pNot = NewNode ( pNode->m_tSpec );
			pNot->m_iOp = MGPU_OP_OR;
			pNot->m_dChildren = dNots;
		}
		pNode->m_iOp = MGPU_OP_ANDNOT;
		pNode->m_iOpArg = 0;
		pNode->m_dChildren.clear();
		pNode->m_dChildren.push_back ( pAnd );
		pNode->m_dChildren.push_back ( pNot );
		return true;
	}

	bool CheckQuorumProximity ( PNode_t * pNode )
	{
		if ( !pNode )
			return true;
		if ( pNode->m_iOp==MGPU_OP_QUORUM && !pNode->m_dWords.empty() )
		{
			if ( !( pNode->m_iOpArg>0 && ( !pNode->m_bPercentOp || pNode->m_iOpArg<=100 ) ) )
			{
				Fail ( MGPU_E_BAD_QUERY, pNode->m_bPercentOp ? "quorum threshold out of bounds 0.0 and 1.0f" : "quorum threshold too low (" + std::to_string ( pNode->m_iOpArg ) + ")" );
				return false;
			}
		}
		if ( pNode->m_iOp==MGPU_OP_PROXIMITY && !pNode->m_dWords.empty() && pNode->m_iOpArg<1 )
		{
			Fail ( MGPU_E_BAD_QUERY, "proximity threshold too low (" + std::to_string ( pNode->m_iOpArg ) + ")" );
			return false;
		}
		for ( PNode_t * p : pNode->m_dChildren )
			if ( !CheckQuorumProximity ( p ) )
				return false;
		return true;
	}

public:
	//////////////////////////////////////////////////////////////////////////
	// sphTransformExtendedQuery, src/sphinx.cpp:15345-15360 (no bigram index, no boolean simplification)

	/// TransformQuorum, src/sphinx.cpp:14643-14669: "a b c"/1 is an OR over its keywords
	void TransformQuorum ( PNode_t * pNode )
	{
		if ( pNode->m_iOp!=MGPU_OP_QUORUM )
		{
			for ( PNode_t * p : pNode->m_dChildren )
				TransformQuorum ( p );
			return;
		}
		if ( pNode->m_iOpArg!=1 )
			return;
		for ( const PWord_t & w : pNode->m_dWords )
		{
			PNode_t * pAnd = NewNode ( pNode->m_tSpec );
			pAnd->m_dWords.push_back ( w );
			pNode->m_dChildren.push_back ( pAnd );
		}
		pNode->m_dWords.clear();
		pNode->m_iOp = MGPU_OP_OR;
		pNode->m_iOpArg = 0;
		pNode->m_bPercentOp = false;
	}

	/// TransformNear, src/sphinx.cpp:15046-15100: (A B) NEAR C is A NEAR B NEAR C
	void TransformNear ( PNode_t * pNode )
	{
		if ( pNode->m_iOp==MGPU_OP_NEAR )
		{
			bool bAgain = true;
			while ( bAgain )
			{
				bAgain = false;
				std::vector<PNode_t*> dOut;
				for ( PNode_t * pChild : pNode->m_dChildren )
					if ( pChild->m_iOp==MGPU_OP_AND && !pChild->m_dChildren.empty() )
					{
						dOut.insert ( dOut.end(), pChild->m_dChildren.begin(), pChild->m_dChildren.end() );
						bAgain = true;
					} else
						dOut.push_back ( pChild );
				pNode->m_dChildren.swap ( dOut );
			}
		}
		for ( PNode_t * p : pNode->m_dChildren )
			TransformNear ( p );
	}

	/// TagExcluded, src/sphinx.cpp:15103-15128
	void TagExcluded ( PNode_t * pNode, bool bNot )
	{
		if ( pNode->m_iOp==MGPU_OP_ANDNOT && pNode->m_dChildren.size()==2 )
		{
			TagExcluded ( pNode->m_dChildren[0], bNot );
			TagExcluded ( pNode->m_dChildren[1], !bNot );
		} else if ( !pNode->m_dChildren.empty() )
		{
			for ( PNode_t * p : pNode->m_dChildren )
				TagExcluded ( p, bNot );
		} else
			for ( PWord_t & w : pNode->m_dWords )
				w.m_bExcluded = bNot;
	}

private:
	/// XQParseHelper_c::FixupTree, :343-387; an empty tree comes back as one node without words or children
	/// XQParseHelper_c::DeleteNodesWOFields, :217-255: children whose field limit matched no field of the schema (@@relaxed) go away
	void DeleteNodesWOFields ( PNode_t * pNode )
	{
		if ( !pNode )
			return;
		for ( size_t i=0; i<pNode->m_dChildren.size(); )
		{
			if ( pNode->m_dChildren[i]->m_tSpec.m_uFieldMask==0 )
			{
				pNode->m_dChildren[i] = pNode->m_dChildren.back();	// RemoveFast
				pNode->m_dChildren.pop_back();
			} else
			{
				DeleteNodesWOFields ( pNode->m_dChildren[i] );
				++i;
			}
		}
	}

	PNode_t * FixupTree ( PNode_t * pRoot )
	{
		if ( !m_bStopOnInvalid )
			DeleteNodesWOFields ( pRoot );
		pRoot = SweepNulls ( pRoot );
		FixupDegenerates ( pRoot );
		FixupNulls ( pRoot );
		if ( !FixupNots ( pRoot ) )
			return nullptr;
		if ( !CheckQuorumProximity ( pRoot ) )
			return nullptr;
		if ( pRoot && pRoot->m_iOp==MGPU_OP_NOT && !pRoot->m_bNullOp && !pRoot->m_iOpArg )
		{
			Fail ( MGPU_E_BAD_QUERY, "query is non-computable (single NOT operator)" );
			return nullptr;
		}
		if ( pRoot && ( pRoot->m_bNullOp || pRoot->m_iOp==MGPU_OP_NOT ) )
			pRoot = nullptr;
		return pRoot ? pRoot : NewNode ( *m_dStateSpec.back() );
	}
};

/// index-mode tokenization of a text for the KEYWORDS command (CSphIndex_VLN::DoGetKeywords, src/sphinx.cpp: Clone ( SPH_CLONE_INDEX ),
/// none of the query specials): { keyword, in-query position }, stop words dropped after taking their position
void TokenizePlain ( const mgpu_parser_settings & s, const char * sText, std::vector<std::pair<std::string,int>> & dOut )
{
	QueryTokenizer_c tTok;
	tTok.m_iMinWordLen = s.min_word_len>0 ? s.min_word_len : 1;
	tTok.m_bCjk = s.ngram_cjk!=0;
	tTok.m_bIndexMode = true;
	std::string sBuf ( sText ? sText : "" );
	tTok.SetBuffer ( sBuf.c_str(), (int)sBuf.size() );
	std::unordered_set<std::string> hStop;
	for ( int i=0; i<s.n_stopwords; ++i )
		if ( s.stopwords[i] )
			hStop.insert ( s.stopwords[i] );
	int iPos = 0;
	while ( const char * sToken = tTok.GetToken() )
	{
		iPos += 1 + tTok.GetOvershortCount()*( s.overshort_step ? 1 : 0 );
		if ( hStop.count ( sToken ) )
		{
			if ( !s.stopword_step )
				--iPos;
			continue;
		}
		dOut.push_back ( { sToken, iPos } );
	}
}

} // namespace mgpu

//////////////////////////////////////////////////////////////////////////
// C ABI
//////////////////////////////////////////////////////////////////////////

struct mgpu_parsed
{
	std::vector<mgpu_xqnode>	m_dNodes;
	std::vector<int32_t>		m_dChildren;
	std::vector<mgpu_xqkeyword>	m_dWords;
	std::vector<std::string>	m_dStrings;
	int							m_iRoot = 0;
	std::string					m_sError, m_sWarning;
	std::string					m_sExplain;		///< SHOW PLAN's transformed_tree
	int							m_iRanker = -1;
};

namespace mgpu
{

static int CountWords ( const PNode_t * p )
{
	int n = (int)p->m_dWords.size();
	for ( const PNode_t * c : p->m_dChildren )
		n += CountWords ( c );
	return n;
}

/// the tree as SHOW PLAN prints it: BuildProfileBson + RenderPlainBsonPlan (src/sphinxsearch.cpp:300-335, 430-495) with the plain
/// indent and line break ("  ", "\n"); a keyword is KEYWORD(word, querypos=N[, excluded][, field_start][, field_end][, boost=%f])
static void Explain ( const PNode_t * p, const std::vector<std::string> & dFields, int iIndent, std::string & sOut )
{
	static const char * dNames[] = { "AND", "OR", "MAYBE", "NOT", "ANDNOT", "BEFORE", "PHRASE", "PROXIMITY", "QUORUM", "NEAR", "NOTNEAR", "SENTENCE", "PARAGRAPH" };
	if ( iIndent )
		sOut += "\n";
	for ( int i=0; i<iIndent; ++i )
		sOut += "  ";
	sOut += ( p->m_iOp>=0 && p->m_iOp<=MGPU_OP_PARAGRAPH ) ? dNames[p->m_iOp] : "OPERATOR";
	sOut += "(";
	bool bComma = false;
	auto fnItem = [&] ( const std::string & s ) { if ( bComma ) sOut += ", "; sOut += s; bComma = true; };
	if ( p->m_iOp==MGPU_OP_PROXIMITY || p->m_iOp==MGPU_OP_NEAR )
		fnItem ( "distance=" + std::to_string ( p->m_iOpArg ) );
	else if ( p->m_iOp==MGPU_OP_QUORUM )
		fnItem ( "count=" + std::to_string ( p->m_iOpArg ) );
	if ( !p->m_dChildren.empty() && !p->m_dWords.empty() )
		fnItem ( "virtually-plain" );
	if ( !p->m_dWords.empty() )
	{
		// AddAccessSpecsBson: only keyword nodes carry their limits into the plan
		if ( p->m_tSpec.m_bFieldSpec && p->m_tSpec.m_uFieldMask!=0xFFFFFFFFu )
		{
			std::string s = "fields=(";
			bool bFirst = true;
			for ( size_t i=0; i<dFields.size() && i<32; ++i )
				if ( p->m_tSpec.m_uFieldMask & ( 1u<<i ) )
				{
					s += ( bFirst ? "" : ", " ) + dFields[i];
					bFirst = false;
				}
			fnItem ( s + ")" );
		}
		if ( p->m_tSpec.m_iFieldMaxPos )
			fnItem ( "max_field_pos=" + std::to_string ( p->m_tSpec.m_iFieldMaxPos ) );
	}
	if ( p->m_dChildren.empty() )
		for ( const PWord_t & w : p->m_dWords )
		{
			std::string s = "KEYWORD(" + w.m_sWord + ", querypos=" + std::to_string ( w.m_iAtomPos );
			if ( w.m_bExcluded )	s += ", excluded";
			if ( w.m_bFieldStart )	s += ", field_start";
			if ( w.m_bFieldEnd )	s += ", field_end";
			if ( w.m_fBoost!=1.0f )
			{
				char sBuf[64];
				snprintf ( sBuf, sizeof(sBuf), ", boost=%f", (double)w.m_fBoost );
				s += sBuf;
			}
			fnItem ( s + ")" );
		}
	else
		for ( const PNode_t * c : p->m_dChildren )
		{
			if ( bComma )
				sOut += ", ";
			bComma = true;
			Explain ( c, dFields, iIndent+1, sOut );
		}
	sOut += ")";
}

static int Flatten ( const PNode_t * p, mgpu_parsed & tOut )
{
	const int iNode = (int)tOut.m_dNodes.size();
	tOut.m_dNodes.emplace_back();
	mgpu_xqnode n;
	memset ( &n, 0, sizeof(n) );
	n.op = p->m_bNullOp ? MGPU_OP_AND : p->m_iOp;
	n.oparg = p->m_iOpArg;
	n.field_mask = p->m_tSpec.m_uFieldMask;
	n.field_max_pos = p->m_tSpec.m_iFieldMaxPos;
	if ( p->m_iOp==MGPU_OP_QUORUM && p->m_bPercentOp )
		n.oparg = (int)floor ( 1.0f / 100.0f * p->m_iOpArg * (int)p->m_dWords.size() + 0.5f );	// ExtQuorum_c::GetThreshold, src/searchnode.cpp:4597-4600
	if ( p->m_iOp!=MGPU_OP_QUORUM && p->m_iOp!=MGPU_OP_PROXIMITY && p->m_iOp!=MGPU_OP_NEAR && p->m_iOp!=MGPU_OP_NOTNEAR )
		n.oparg = 0;	// elsewhere m_iOpArg is the parser's own "nulls were swept" flag
	n.first_word = (int)tOut.m_dWords.size();
	n.n_words = (int)p->m_dWords.size();
	for ( const PWord_t & w : p->m_dWords )
	{
		mgpu_xqkeyword k;
		memset ( &k, 0, sizeof(k) );
		k.word = (const char *)(intptr_t)tOut.m_dStrings.size();	// index now, pointer once the strings stop moving
		tOut.m_dStrings.push_back ( w.m_sWord );
		k.atom_pos = w.m_iAtomPos;
		k.boost = w.m_fBoost;
		k.field_start = w.m_bFieldStart;
		k.field_end = w.m_bFieldEnd;
		k.excluded = w.m_bExcluded;
		tOut.m_dWords.push_back ( k );
	}
	std::vector<int> dKids;
	for ( const PNode_t * c : p->m_dChildren )
		dKids.push_back ( Flatten ( c, tOut ) );
	n.first_child = (int)tOut.m_dChildren.size();
	n.n_children = (int)dKids.size();
	for ( int k : dKids )
		tOut.m_dChildren.push_back ( k );
	tOut.m_dNodes[iNode] = n;
	return iNode;
}

} // namespace mgpu

extern "C"
{

static int ParseQueryImpl ( const mgpu_parser_settings * settings, const char * text, mgpu_parsed ** out );

int mgpu_parse_query ( const mgpu_parser_settings * settings, const char * text, mgpu_parsed ** out )
{
	// no exception crosses the ABI
	try
	{
		return ParseQueryImpl ( settings, text, out );
	} catch ( ... )
	{
		if ( out )
			*out = nullptr;
		return MGPU_E_NOMEM;
	}
}

static int ParseQueryImpl ( const mgpu_parser_settings * settings, const char * text, mgpu_parsed ** out )
{
	if ( !settings || !out || settings->n_fields<0 || ( settings->n_fields && !settings->field_names ) || ( settings->n_stopwords>0 && !settings->stopwords ) )
		return MGPU_E_BAD_QUERY;
	std::unique_ptr<mgpu_parsed> pRes ( new mgpu_parsed );
	std::string sQuery ( text ? text : "" );

	// legacy match modes, PrepareQueryEmulation (src/searchd.cpp:2141-2190): escape the syntax, wrap for any / phrase, pick the ranker
	const int iMode = settings->match_mode;
	if ( iMode==MGPU_MATCH_ALL || iMode==MGPU_MATCH_ANY || iMode==MGPU_MATCH_PHRASE )
	{
		std::string s;
		if ( iMode!=MGPU_MATCH_ALL )
			s += '"';
		for ( char c : sQuery )
		{
			if ( strchr ( "<\\()|-!@~\"&/^$=", c ) && c )
				s += '\\';
			s += c;
		}
		if ( iMode==MGPU_MATCH_ANY )
			s += "\"/1";
		else if ( iMode==MGPU_MATCH_PHRASE )
			s += '"';
		sQuery.swap ( s );
		pRes->m_iRanker = ( iMode==MGPU_MATCH_ANY ) ? MGPU_RANK_MATCHANY : MGPU_RANK_PROXIMITY;
	} else if ( iMode==MGPU_MATCH_BOOLEAN )
		pRes->m_iRanker = MGPU_RANK_NONE;
	else if ( iMode!=MGPU_MATCH_EXTENDED )
		return MGPU_E_BAD_QUERY;

	mgpu::Parser_c tParser ( *settings );
	mgpu::PNode_t * pRoot = tParser.Parse ( sQuery.c_str() );
	pRes->m_sWarning = tParser.m_sWarning;
	int iRes = MGPU_OK;
	if ( !pRoot )
	{
		pRes->m_sError = tParser.m_sError.empty() ? "parse error" : tParser.m_sError;
		iRes = tParser.m_iErrorCode!=MGPU_OK ? tParser.m_iErrorCode : MGPU_E_BAD_QUERY;
	} else
	{
		tParser.TransformQuorum ( pRoot );
		tParser.TransformNear ( pRoot );
		tParser.TagExcluded ( pRoot, false );
		{
			std::vector<std::string> dFields;
			for ( int i=0; i<settings->n_fields; ++i )
				dFields.push_back ( settings->field_names[i] ? settings->field_names[i] : "" );
			mgpu::Explain ( pRoot, dFields, 0, pRes->m_sExplain );
		}
		pRes->m_dStrings.reserve ( mgpu::CountWords ( pRoot ) );
		pRes->m_iRoot = mgpu::Flatten ( pRoot, *pRes );
		for ( auto & k : pRes->m_dWords )
			k.word = pRes->m_dStrings[(size_t)(intptr_t)k.word].c_str();
	}
	*out = pRes.release();
	return iRes;
}

int mgpu_parsed_fill ( const mgpu_parsed * p, mgpu_query * q )
{
	if ( !p || !q || !p->m_sError.empty() )
		return MGPU_E_BAD_QUERY;
	q->nodes = p->m_dNodes.data();
	q->n_nodes = (int32_t)p->m_dNodes.size();
	q->root = p->m_iRoot;
	q->children = p->m_dChildren.data();
	q->n_children = (int32_t)p->m_dChildren.size();
	q->words = p->m_dWords.data();
	q->n_words = (int32_t)p->m_dWords.size();
	if ( p->m_iRanker>=0 )
		q->ranker = p->m_iRanker;
	return MGPU_OK;
}

const char * mgpu_parsed_error ( const mgpu_parsed * p )		{ return p ? p->m_sError.c_str() : ""; }
const char * mgpu_parsed_warning ( const mgpu_parsed * p )	{ return p ? p->m_sWarning.c_str() : ""; }
const char * mgpu_parsed_explain ( const mgpu_parsed * p )	{ return p ? p->m_sExplain.c_str() : ""; }
void mgpu_parsed_free ( mgpu_parsed * p )					{ delete p; }

} // extern "C"
