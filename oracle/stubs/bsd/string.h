/* stand-in for libbsd's <bsd/string.h>: glibc >= 2.38 already provides strlcpy/strlcat */
#include <string.h>
