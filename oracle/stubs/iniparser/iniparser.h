/* declaration-only stub: lets the reference radio.c compile for the oracle; nothing here is ever called */
#ifndef STUB_INIPARSER_H
#define STUB_INIPARSER_H
#include <stdio.h>
typedef struct _dictionary_ dictionary;
dictionary *iniparser_load(const char *);
void iniparser_freedict(dictionary *);
int iniparser_getnsec(const dictionary *);
const char *iniparser_getsecname(const dictionary *, int);
const char *iniparser_getstring(const dictionary *, const char *, const char *);
int iniparser_getint(const dictionary *, const char *, int);
double iniparser_getdouble(const dictionary *, const char *, double);
int iniparser_getboolean(const dictionary *, const char *, int);
int iniparser_find_entry(const dictionary *, const char *);
int iniparser_getsecnkeys(const dictionary *, const char *);
const char **iniparser_getseckeys(const dictionary *, const char *, const char **);
void iniparser_dump_ini(const dictionary *, FILE *);
#endif
