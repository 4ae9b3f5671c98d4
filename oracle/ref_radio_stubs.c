/* oracle/ref_radio_stubs.c -- link-time stand-ins for the parts of radiod that radio.c references but the oracle
 * never runs (config parser, multicast, status protocol, demodulators, avahi, opus).  TEST INFRASTRUCTURE.
 * Each aborts if it is ever reached: the oracle only calls downconvert() and what it uses (filter.c, osc.c, misc.c). */
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#define STUB(name)                                                     \
  void name(void) {                                                    \
    fprintf(stderr, "oracle/_ref: unexpected call of %s\n", #name);    \
    abort();                                                           \
  }
STUB(avahi_start) STUB(config2_getstring) STUB(config_getboolean) STUB(config_getdouble) STUB(config_getint)
STUB(config_getstring) STUB(config_validate) STUB(config_validate_section) STUB(decode_radio_commands) STUB(demod_fm)
STUB(demod_linear) STUB(demod_name_from_type) STUB(demod_spectrum) STUB(demod_wfm) STUB(encoding_string) STUB(formatsock)
STUB(gen_sdes) STUB(gen_sr) STUB(iniparser_freedict) STUB(iniparser_getnsec) STUB(iniparser_getsecname) STUB(iniparser_load)
STUB(join_group) STUB(listen_mcast) STUB(loadpreset) STUB(make_maddr) STUB(opus_encoder_destroy) STUB(output_mcast)
STUB(radio_status) STUB(resolve_mcast) STUB(send_radio_status) STUB(set_defaults) STUB(setport)
/* data the headers declare extern (radio.h:363, multicast.h:16, avahi.h:10, main.c) */
char const *Channel_keys[] = {NULL};
char const *Default_mcast_iface = NULL;
char const *Name = "ka9q-oracle";
bool Static_avahi = false;
