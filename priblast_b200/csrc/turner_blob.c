/* Embeds priblast_b200/data/turner99.bin (see turner_params.h) into the object file.
 * Compile with -DPRIB_TURNER_BIN='"<absolute path to turner99.bin>"'. */
#include "turner_params.h"

#ifndef PRIB_TURNER_BIN
#error "define PRIB_TURNER_BIN to the absolute path of turner99.bin"
#endif

__asm__(".section .rodata\n"
        ".balign 16\n"
        ".global prib_turner_blob_begin\n"
        "prib_turner_blob_begin:\n"
        ".incbin \"" PRIB_TURNER_BIN "\"\n"
        ".global prib_turner_blob_end\n"
        "prib_turner_blob_end:\n"
        ".byte 0\n"
        ".previous\n");

extern const unsigned char prib_turner_blob_begin[];
extern const unsigned char prib_turner_blob_end[];

const prib_turner_params *prib_turner_embedded(void) {
  if ((unsigned long)(prib_turner_blob_end - prib_turner_blob_begin) != sizeof(prib_turner_params)) return 0;
  const prib_turner_params *p = (const prib_turner_params *)prib_turner_blob_begin;
  if (p->magic != PRIB_TURNER_MAGIC || p->version != PRIB_TURNER_VERSION) return 0;
  return p;
}
