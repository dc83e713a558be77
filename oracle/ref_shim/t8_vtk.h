/* t8mini forwarding header (TEST INFRASTRUCTURE, see t8.h) */
#include <t8.h>
