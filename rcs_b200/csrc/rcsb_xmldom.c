/*
 * Element-only XML DOM on top of expat, with XInclude splicing, plus attribute parsing
 * with the rules of the reference's helpers (src/RcsCore/Rcs_parser.c:320-740):
 *  - tags are matched by name, so character data and comments are not kept;
 *  - <xi:include href="f"/> is replaced in place by the root element of f, f resolved
 *    relative to the including file (what libxml2's xmlXIncludeProcess does for the
 *    reference, src/RcsCore/Rcs_parser.c:140-180);
 *  - number lists are split at spaces and converted with the "C" locale; an attribute
 *    whose count of numbers differs from the requested count is ignored.
 */
#include "rcsb_internal.h"

#include <expat.h>
#include <locale.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

typedef struct
{
  RcsbXmlNode* root;
  RcsbXmlNode* cur;
  const char* file;
  int depth;
  int skipDepth;
} DomBuilder;

static RcsbXmlNode* parseFileDepth(const char* fileName, int depth);

static char* strDup(const char* s)
{
  size_t n = strlen(s) + 1;
  char* d = (char*) malloc(n);
  memcpy(d, s, n);
  return d;
}

static void linkChild(RcsbXmlNode* parent, RcsbXmlNode* child)
{
  child->parent = parent;
  child->next = NULL;
  if (parent->last)
  {
    parent->last->next = child;
  }
  else
  {
    parent->children = child;
  }
  parent->last = child;
}

static bool isXInclude(const char* tag)
{
  const char* colon = strrchr(tag, ':');
  return colon && strcmp(colon + 1, "include") == 0;
}

static char* siblingPath(const char* baseFile, const char* href)
{
  const char* slash = baseFile ? strrchr(baseFile, '/') : NULL;
  if (href[0] == '/' || !slash)
  {
    return strDup(href);
  }
  size_t dirLen = (size_t)(slash - baseFile) + 1;
  char* p = (char*) malloc(dirLen + strlen(href) + 1);
  memcpy(p, baseFile, dirLen);
  strcpy(p + dirLen, href);
  return p;
}

static void XMLCALL domStart(void* ud, const XML_Char* tag, const XML_Char** atts)
{
  DomBuilder* b = (DomBuilder*) ud;

  if (b->skipDepth > 0)
  {
    b->skipDepth++;
    return;
  }

  if (isXInclude(tag))
  {
    b->skipDepth = 1;   /* ignore xi:fallback children */
    const char* href = NULL;
    for (int i = 0; atts[i]; i += 2)
    {
      if (strcmp(atts[i], "href") == 0)
      {
        href = atts[i + 1];
      }
    }
    if (!href || b->depth > 32)
    {
      return;
    }
    char* path = siblingPath(b->file, href);
    RcsbXmlNode* inc = parseFileDepth(path, b->depth + 1);
    if (!inc)
    {
      RCSB_LOG(1, "xi:include of \"%s\" failed", path);
    }
    else if (b->cur)
    {
      linkChild(b->cur, inc);
    }
    else if (!b->root)
    {
      b->root = inc;
    }
    free(path);
    return;
  }

  RcsbXmlNode* n = (RcsbXmlNode*) calloc(1, sizeof(RcsbXmlNode));
  n->name = strDup(tag);
  RcsbXmlAttr* tail = NULL;
  for (int i = 0; atts[i]; i += 2)
  {
    RcsbXmlAttr* a = (RcsbXmlAttr*) calloc(1, sizeof(RcsbXmlAttr));
    a->name = strDup(atts[i]);
    a->value = strDup(atts[i + 1]);
    if (tail)
    {
      tail->next = a;
    }
    else
    {
      n->attrs = a;
    }
    tail = a;
  }

  if (b->cur)
  {
    linkChild(b->cur, n);
  }
  else if (!b->root)
  {
    b->root = n;
  }
  b->cur = n;
}

static void XMLCALL domEnd(void* ud, const XML_Char* tag)
{
  DomBuilder* b = (DomBuilder*) ud;
  (void) tag;
  if (b->skipDepth > 0)
  {
    b->skipDepth--;
    return;
  }
  if (b->cur)
  {
    b->cur = b->cur->parent;
  }
}

static RcsbXmlNode* parseBufferDepth(const char* buf, size_t size, const char* file, int depth)
{
  DomBuilder b;
  memset(&b, 0, sizeof(b));
  b.file = file;
  b.depth = depth;

  XML_Parser p = XML_ParserCreate(NULL);
  XML_SetUserData(p, &b);
  XML_SetElementHandler(p, domStart, domEnd);

  if (XML_Parse(p, buf, (int) size, 1) == XML_STATUS_ERROR)
  {
    rcsb_set_error("XML error in %s line %lu: %s", file ? file : "<buffer>",
                   (unsigned long) XML_GetCurrentLineNumber(p),
                   XML_ErrorString(XML_GetErrorCode(p)));
    RCSB_LOG(1, "%s", rcsb_error_string());
    XML_ParserFree(p);
    rcsb_xml_free(b.root);
    return NULL;
  }
  XML_ParserFree(p);
  return b.root;
}

static RcsbXmlNode* parseFileDepth(const char* fileName, int depth)
{
  FILE* f = fopen(fileName, "rb");
  if (!f)
  {
    return NULL;
  }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  char* buf = (char*) malloc((size_t) sz + 1);
  size_t got = fread(buf, 1, (size_t) sz, f);
  fclose(f);
  RcsbXmlNode* root = parseBufferDepth(buf, got, fileName, depth);
  free(buf);
  return root;
}

RcsbXmlNode* rcsb_xml_parse_file(const char* fileName)
{
  return parseFileDepth(fileName, 0);
}

RcsbXmlNode* rcsb_xml_parse_buffer(const char* buf, size_t size)
{
  /* tolerate a trailing '\0' counted in size, as RcsGraph_create passes strlen+1 */
  while (size > 0 && buf[size - 1] == '\0')
  {
    size--;
  }
  return parseBufferDepth(buf, size, NULL, 0);
}

void rcsb_xml_free(RcsbXmlNode* n)
{
  while (n)
  {
    RcsbXmlNode* next = n->next;
    rcsb_xml_free(n->children);
    for (RcsbXmlAttr* a = n->attrs; a;)
    {
      RcsbXmlAttr* an = a->next;
      free(a->name);
      free(a->value);
      free(a);
      a = an;
    }
    free(n->name);
    free(n);
    n = next;
  }
}

const char* rcsb_xml_attr(const RcsbXmlNode* n, const char* name)
{
  if (!n || !name)
  {
    return NULL;
  }
  for (const RcsbXmlAttr* a = n->attrs; a; a = a->next)
  {
    if (strcmp(a->name, name) == 0)
    {
      return a->value;
    }
  }
  return NULL;
}

/* number of space-separated tokens (src/RcsCore/Rcs_utils.c:350) */
unsigned int rcsb_xml_num_strings(const RcsbXmlNode* n, const char* tag)
{
  const char* v = rcsb_xml_attr(n, tag);
  if (!v)
  {
    return 0;
  }
  unsigned int count = 0;
  const char* p = v;
  while (*p)
  {
    while (*p == ' ')
    {
      p++;
    }
    if (!*p)
    {
      break;
    }
    count++;
    while (*p && *p != ' ')
    {
      p++;
    }
  }
  return count;
}

/* src/RcsCore/Rcs_utils.c:466 String_toDoubleArray_l: all-or-nothing */
bool rcsb_xml_vecn(const RcsbXmlNode* n, const char* tag, double* x, unsigned int cnt)
{
  const char* v = rcsb_xml_attr(n, tag);
  if (!v || !*v || cnt == 0 || !x)
  {
    return false;
  }

  static locale_t cLocale = (locale_t) 0;
  if (!cLocale)
  {
    cLocale = newlocale(LC_NUMERIC_MASK, "C", (locale_t) 0);
  }

  double* tmp = (double*) calloc(cnt, sizeof(double));
  unsigned int matched = 0;
  const char* p = v;
  bool ok = true;

  while (*p && ok)
  {
    while (*p == ' ')
    {
      p++;
    }
    if (!*p)
    {
      break;
    }
    /* the reference sscanf("%255s")s each space-delimited token, which also strips
     * other leading white space, then strtod_l's it */
    char tok[256];
    size_t len = 0;
    while (p[len] && p[len] != ' ' && len < 255)
    {
      tok[len] = p[len];
      len++;
    }
    tok[len] = '\0';
    p += len;
    while (*p && *p != ' ')
    {
      p++;
    }

    const char* t = tok;
    while (*t == '\t' || *t == '\n' || *t == '\r')
    {
      t++;
    }
    if (!*t)
    {
      continue;   /* sscanf("%s") fails on an all-white-space token: not counted */
    }

    double val = strtod_l(t, NULL, cLocale);
    if (!isfinite(val))
    {
      ok = false;
      break;
    }
    if (matched < cnt)
    {
      tmp[matched] = val;
    }
    matched++;
  }

  if (ok && matched == cnt)
  {
    memcpy(x, tmp, cnt * sizeof(double));
  }
  else
  {
    ok = false;
    RCSB_LOG(4, "[%s]: %u values parsed, %u expected", v, matched, cnt);
  }
  free(tmp);
  return ok;
}

/* only a case-insensitive "true" is true (src/RcsCore/Rcs_utils.c:334) */
bool rcsb_xml_bool(const RcsbXmlNode* n, const char* tag, bool* x)
{
  const char* v = rcsb_xml_attr(n, tag);
  if (!v)
  {
    return false;
  }
  if (x)
  {
    *x = (strcasecmp(v, "true") == 0);
  }
  return true;
}

/* 6 numbers: x y z + x-y-z Euler angles in degrees; 12: x y z + row-major rotation
 * (src/RcsCore/Rcs_parser.c:702-740) */
bool rcsb_xml_htr(const RcsbXmlNode* n, const char* tag, HTr* A)
{
  double x[12];
  unsigned int nItems = rcsb_xml_num_strings(n, tag);
  if (nItems != 6 && nItems != 12)
  {
    return false;
  }
  if (!rcsb_xml_vecn(n, tag, x, nItems))
  {
    return false;
  }
  memcpy(A->org, x, 3 * sizeof(double));
  if (nItems == 6)
  {
    for (int i = 3; i < 6; ++i)
    {
      x[i] *= (M_PI / 180.0);
    }
    Mat3d_fromEulerAngles(A->rot, &x[3]);
  }
  else
  {
    memcpy(A->rot, &x[3], 9 * sizeof(double));
  }
  return true;
}

/* returns strlen+1 like getXMLNodePropertyStringN, 0 if the attribute is absent */
unsigned int rcsb_xml_string(const RcsbXmlNode* n, const char* tag, char* str, unsigned int cap)
{
  const char* v = rcsb_xml_attr(n, tag);
  if (!v)
  {
    return 0;
  }
  if (str && cap > 0)
  {
    strncpy(str, v, cap);
    str[cap - 1] = '\0';
  }
  return (unsigned int) strlen(v) + 1;
}
