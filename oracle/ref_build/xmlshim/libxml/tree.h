/*
 * Minimal stand-in for <libxml/tree.h>, backed by expat (oracle/ref_build/xmlshim/xmlshim.c).
 *
 * TEST INFRASTRUCTURE ONLY. libxml2 is not installed in this image, expat is. The
 * unmodified reference sources under /root/reference are compiled against this header so
 * that the reference's own loader and math can run here as the parity anchor
 * (oracle/_ref/librcsref.so). Nothing in the product (rcs_b200/) includes this file.
 *
 * Only what the reference's hot-path loader touches is provided: an element-only DOM
 * (the reference matches children by tag name, so text/comment nodes are not needed),
 * attribute lookup, and xi:include splicing.
 */
#ifndef ORACLE_XMLSHIM_TREE_H
#define ORACLE_XMLSHIM_TREE_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char xmlChar;
#define BAD_CAST (xmlChar*)
#define LIBXML_TEST_VERSION

typedef enum
{
  XML_ELEMENT_NODE = 1
} xmlElementType;

typedef enum
{
  XML_PARSE_NOERROR = 1 << 5,
  XML_PARSE_NOWARNING = 1 << 6,
  XML_PARSE_XINCLUDE = 1 << 10
} xmlParserOption;

typedef struct _xmlAttr
{
  xmlChar* name;
  xmlChar* value;
  struct _xmlAttr* next;
} xmlAttr;

struct _xmlDoc;

typedef struct _xmlNode
{
  void* _private;
  xmlElementType type;
  const xmlChar* name;
  struct _xmlNode* children;
  struct _xmlNode* last;
  struct _xmlNode* parent;
  struct _xmlNode* next;
  struct _xmlNode* prev;
  struct _xmlDoc* doc;
  xmlAttr* properties;
  xmlChar* content;   /* concatenated character data of the element */
  char* baseFile;     /* file this element was read from (xi:include resolution) */
} xmlNode;
typedef xmlNode* xmlNodePtr;

typedef struct _xmlDoc
{
  xmlNode* root;
  char* url;
} xmlDoc;
typedef xmlDoc* xmlDocPtr;

typedef void (*xmlFreeFunc)(void*);
extern xmlFreeFunc xmlFree;

xmlDocPtr xmlReadFile(const char* filename, const char* encoding, int options);
xmlDocPtr xmlParseFile(const char* filename);
xmlDocPtr xmlReadMemory(const char* buffer, int size, const char* url,
                        const char* encoding, int options);
int xmlXIncludeProcess(xmlDocPtr doc);
xmlNodePtr xmlDocGetRootElement(xmlDocPtr doc);
void xmlFreeDoc(xmlDocPtr doc);
void xmlCleanupParser(void);

xmlChar* xmlGetProp(const xmlNode* node, const xmlChar* name);
xmlChar* xmlNodeGetContent(const xmlNode* node);
int xmlStrcmp(const xmlChar* a, const xmlChar* b);
int xmlStrcasecmp(const xmlChar* a, const xmlChar* b);

void xmlDocDumpMemory(xmlDocPtr doc, xmlChar** mem, int* size);
int xmlDocFormatDump(FILE* f, xmlDocPtr doc, int format);

#ifdef __cplusplus
}
#endif

#endif
