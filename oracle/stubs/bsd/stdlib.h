/* stand-in for libbsd's <bsd/stdlib.h> */
#include <stdlib.h>
