/*
 * expat-backed implementation of the libxml2 subset declared in libxml/tree.h.
 *
 * TEST INFRASTRUCTURE ONLY (see libxml/tree.h). Builds an element-only DOM and splices
 * <xi:include href="..."/> elements in place at parse time: the included document's root
 * element takes the position of the include element, with hrefs resolved relative to the
 * including file, which is what libxml2's XInclude processing does for the graph files
 * under /root/reference/config/xml.
 */
#include "libxml/tree.h"

#include <expat.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

static void shimFree(void* p)
{
  free(p);
}
xmlFreeFunc xmlFree = shimFree;

typedef struct
{
  xmlDoc* doc;
  xmlNode* cur;       /* element currently open */
  const char* file;   /* file being parsed (may be NULL for memory buffers) */
  int depth;          /* include nesting depth */
  int skip;           /* >0 while inside an xi:include element (fallback children) */
  int failed;
} ParseCtx;

static xmlDoc* parseFileInto(const char* filename, int depth);

static char* dupStr(const char* s)
{
  size_t n = strlen(s) + 1;
  char* d = (char*) malloc(n);
  memcpy(d, s, n);
  return d;
}

static void appendChild(xmlNode* parent, xmlNode* child)
{
  child->parent = parent;
  child->next = NULL;
  child->prev = parent->last;
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

static void setDocRecursive(xmlNode* n, xmlDoc* doc)
{
  for (; n; n = n->next)
  {
    n->doc = doc;
    setDocRecursive(n->children, doc);
  }
}

static int isIncludeTag(const char* name)
{
  const char* colon = strrchr(name, ':');
  const char* local = colon ? colon + 1 : name;
  return (colon != NULL) && (strcmp(local, "include") == 0);
}

static char* resolveHref(const char* baseFile, const char* href)
{
  if (href[0] == '/' || baseFile == NULL)
  {
    return dupStr(href);
  }
  const char* slash = strrchr(baseFile, '/');
  if (!slash)
  {
    return dupStr(href);
  }
  size_t dirLen = (size_t)(slash - baseFile) + 1;
  char* out = (char*) malloc(dirLen + strlen(href) + 1);
  memcpy(out, baseFile, dirLen);
  strcpy(out + dirLen, href);
  return out;
}

static void XMLCALL onStart(void* ud, const XML_Char* name, const XML_Char** atts)
{
  ParseCtx* ctx = (ParseCtx*) ud;

  if (ctx->skip > 0)
  {
    ctx->skip++;
    return;
  }

  if (isIncludeTag(name))
  {
    const char* href = NULL;
    for (int i = 0; atts[i]; i += 2)
    {
      if (strcmp(atts[i], "href") == 0)
      {
        href = atts[i + 1];
      }
    }
    ctx->skip = 1;
    if (href && ctx->depth < 32)
    {
      char* path = resolveHref(ctx->file, href);
      xmlDoc* inc = parseFileInto(path, ctx->depth + 1);
      if (inc && inc->root)
      {
        xmlNode* r = inc->root;
        inc->root = NULL;
        if (ctx->cur)
        {
          appendChild(ctx->cur, r);
        }
        else
        {
          ctx->doc->root = r;
        }
        setDocRecursive(r, ctx->doc);
        r->next = NULL;
      }
      else
      {
        fprintf(stderr, "[xmlshim] xi:include of \"%s\" failed\n", path);
      }
      if (inc)
      {
        xmlFreeDoc(inc);
      }
      free(path);
    }
    return;
  }

  xmlNode* n = (xmlNode*) calloc(1, sizeof(xmlNode));
  n->type = XML_ELEMENT_NODE;
  n->name = (const xmlChar*) dupStr(name);
  n->doc = ctx->doc;
  n->baseFile = ctx->file ? dupStr(ctx->file) : NULL;

  xmlAttr* tail = NULL;
  for (int i = 0; atts[i]; i += 2)
  {
    xmlAttr* a = (xmlAttr*) calloc(1, sizeof(xmlAttr));
    a->name = (xmlChar*) dupStr(atts[i]);
    a->value = (xmlChar*) dupStr(atts[i + 1]);
    if (tail)
    {
      tail->next = a;
    }
    else
    {
      n->properties = a;
    }
    tail = a;
  }

  if (ctx->cur)
  {
    appendChild(ctx->cur, n);
  }
  else if (!ctx->doc->root)
  {
    ctx->doc->root = n;
  }
  ctx->cur = n;
}

static void XMLCALL onEnd(void* ud, const XML_Char* name)
{
  ParseCtx* ctx = (ParseCtx*) ud;
  (void) name;

  if (ctx->skip > 0)
  {
    ctx->skip--;
    return;
  }
  if (ctx->cur)
  {
    ctx->cur = ctx->cur->parent;
  }
}

static void XMLCALL onText(void* ud, const XML_Char* s, int len)
{
  ParseCtx* ctx = (ParseCtx*) ud;
  if (ctx->skip > 0 || !ctx->cur || len <= 0)
  {
    return;
  }
  xmlNode* n = ctx->cur;
  size_t old = n->content ? strlen((char*) n->content) : 0;
  n->content = (xmlChar*) realloc(n->content, old + (size_t) len + 1);
  memcpy(n->content + old, s, (size_t) len);
  n->content[old + len] = '\0';
}

static xmlDoc* parseBuffer(const char* buf, size_t size, const char* file, int depth)
{
  xmlDoc* doc = (xmlDoc*) calloc(1, sizeof(xmlDoc));
  doc->url = file ? dupStr(file) : NULL;

  ParseCtx ctx;
  memset(&ctx, 0, sizeof(ctx));
  ctx.doc = doc;
  ctx.file = file;
  ctx.depth = depth;

  XML_Parser p = XML_ParserCreate(NULL);
  XML_SetUserData(p, &ctx);
  XML_SetElementHandler(p, onStart, onEnd);
  XML_SetCharacterDataHandler(p, onText);

  if (XML_Parse(p, buf, (int) size, 1) == XML_STATUS_ERROR)
  {
    fprintf(stderr, "[xmlshim] %s:%lu: %s\n", file ? file : "<memory>",
            (unsigned long) XML_GetCurrentLineNumber(p),
            XML_ErrorString(XML_GetErrorCode(p)));
    XML_ParserFree(p);
    xmlFreeDoc(doc);
    return NULL;
  }
  XML_ParserFree(p);
  return doc;
}

static xmlDoc* parseFileInto(const char* filename, int depth)
{
  FILE* f = fopen(filename, "rb");
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
  buf[got] = '\0';
  xmlDoc* doc = parseBuffer(buf, got, filename, depth);
  free(buf);
  return doc;
}

xmlDocPtr xmlReadFile(const char* filename, const char* encoding, int options)
{
  (void) encoding;
  (void) options;
  return parseFileInto(filename, 0);
}

xmlDocPtr xmlParseFile(const char* filename)
{
  return parseFileInto(filename, 0);
}

xmlDocPtr xmlReadMemory(const char* buffer, int size, const char* url,
                        const char* encoding, int options)
{
  (void) url;
  (void) encoding;
  (void) options;
  return parseBuffer(buffer, (size_t) size, NULL, 0);
}

int xmlXIncludeProcess(xmlDocPtr doc)
{
  (void) doc;   /* includes are spliced while parsing */
  return 0;
}

xmlNodePtr xmlDocGetRootElement(xmlDocPtr doc)
{
  return doc ? doc->root : NULL;
}

static void freeNode(xmlNode* n)
{
  while (n)
  {
    xmlNode* next = n->next;
    freeNode(n->children);
    xmlAttr* a = n->properties;
    while (a)
    {
      xmlAttr* an = a->next;
      free(a->name);
      free(a->value);
      free(a);
      a = an;
    }
    free((void*) n->name);
    free(n->content);
    free(n->baseFile);
    free(n);
    n = next;
  }
}

void xmlFreeDoc(xmlDocPtr doc)
{
  if (!doc)
  {
    return;
  }
  freeNode(doc->root);
  free(doc->url);
  free(doc);
}

void xmlCleanupParser(void)
{
}

xmlChar* xmlGetProp(const xmlNode* node, const xmlChar* name)
{
  if (!node || !name)
  {
    return NULL;
  }
  for (const xmlAttr* a = node->properties; a; a = a->next)
  {
    if (strcmp((const char*) a->name, (const char*) name) == 0)
    {
      return (xmlChar*) dupStr((const char*) a->value);
    }
  }
  return NULL;
}

xmlChar* xmlNodeGetContent(const xmlNode* node)
{
  if (!node)
  {
    return NULL;
  }
  return (xmlChar*) dupStr(node->content ? (const char*) node->content : "");
}

int xmlStrcmp(const xmlChar* a, const xmlChar* b)
{
  if (a == b)
  {
    return 0;
  }
  if (!a)
  {
    return -1;
  }
  if (!b)
  {
    return 1;
  }
  return strcmp((const char*) a, (const char*) b);
}

int xmlStrcasecmp(const xmlChar* a, const xmlChar* b)
{
  if (a == b)
  {
    return 0;
  }
  if (!a)
  {
    return -1;
  }
  if (!b)
  {
    return 1;
  }
  return strcasecmp((const char*) a, (const char*) b);
}

static void dumpNode(FILE* f, const xmlNode* n, int indent)
{
  for (; n; n = n->next)
  {
    fprintf(f, "%*s<%s", indent, "", (const char*) n->name);
    for (const xmlAttr* a = n->properties; a; a = a->next)
    {
      fprintf(f, " %s=\"%s\"", (const char*) a->name, (const char*) a->value);
    }
    if (n->children)
    {
      fprintf(f, ">\n");
      dumpNode(f, n->children, indent + 2);
      fprintf(f, "%*s</%s>\n", indent, "", (const char*) n->name);
    }
    else
    {
      fprintf(f, "/>\n");
    }
  }
}

int xmlDocFormatDump(FILE* f, xmlDocPtr doc, int format)
{
  (void) format;
  if (!doc || !f)
  {
    return -1;
  }
  fprintf(f, "<?xml version=\"1.0\"?>\n");
  dumpNode(f, doc->root, 0);
  return 0;
}

void xmlDocDumpMemory(xmlDocPtr doc, xmlChar** mem, int* size)
{
  char* buf = NULL;
  size_t len = 0;
  FILE* f = open_memstream(&buf, &len);
  xmlDocFormatDump(f, doc, 1);
  fclose(f);
  *mem = (xmlChar*) buf;
  *size = (int) len;
}
