/* linne_stdint.h -- fixed-width integer types for the LINNE API.
 * Replaces reference include/linne_stdint.h:1-11 (which also just pulls in <stdint.h>). */
#ifndef LINNE_STDINT_H_INCLUDED
#define LINNE_STDINT_H_INCLUDED
#include <stdint.h>
#endif
