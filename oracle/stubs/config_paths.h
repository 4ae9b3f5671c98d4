/* stand-in for the header the reference Makefile generates (src/Makefile:372-387) */
#ifndef KA9Q_ORACLE_CONFIG_PATHS_H
#define KA9Q_ORACLE_CONFIG_PATHS_H 1
#define CONFDIR "/nonexistent/ka9q-oracle/conf"
#define STATEDIR "/nonexistent/ka9q-oracle/state"
#define PKGDATADIR "/nonexistent/ka9q-oracle/share"
#define PKGLIBDIR "/nonexistent/ka9q-oracle/lib"
#define GIT_HASH "4e0033b4536e625d4ccf962b5a428dcf90f9618c"
#define GIT_TIME "n/a"
#define GIT_BRANCH "n/a"
#define GIT_SUMMARY "oracle build of the reference filter path"
#define GIT_VERSION "oracle"
#define GIT_REMOTE_URL "n/a"
#endif
