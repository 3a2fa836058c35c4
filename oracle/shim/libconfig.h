/*
 * TEST INFRASTRUCTURE ONLY. Stand-in for <libconfig.h> (hyperrealm libconfig, an un-vendored dependency of the reference's
 * control plane: CMakeLists.txt:54-57) with just the calls src/server_config.c makes, so that the reference's own
 * server_config.c and TCP server can be compiled in place and linked against the product library for the integration tests
 * (oracle/Makefile server-tests). Understands the flat `name = value` files of test/resources/*.conf: one setting per line,
 * optional trailing ';', strings in double quotes, integers, floats, '#' and '//' comments.
 */
#ifndef SDRM_LIBCONFIG_SHIM_H
#define SDRM_LIBCONFIG_SHIM_H

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CONFIG_TRUE 1
#define CONFIG_FALSE 0

typedef struct config_setting_t {
    char *name;
    char *text; /* the value with quotes removed */
    int is_string;
} config_setting_t;

typedef struct config_t {
    config_setting_t *settings;
    size_t count;
    char error[160];
} config_t;

static inline void config_init(config_t *config) { memset(config, 0, sizeof(*config)); }

static inline void config_destroy(config_t *config) {
    for (size_t i = 0; i < config->count; i++) {
        free(config->settings[i].name);
        free(config->settings[i].text);
    }
    free(config->settings);
    memset(config, 0, sizeof(*config));
}

static inline const char *config_error_text(const config_t *config) { return config->error; }

static inline char *sdrm_cfg_copy(const char *begin, const char *end) {
    while (begin < end && isspace((unsigned char) *begin)) begin++;
    while (end > begin && isspace((unsigned char) end[-1])) end--;
    char *out = malloc((size_t) (end - begin) + 1);
    memcpy(out, begin, (size_t) (end - begin));
    out[end - begin] = '\0';
    return out;
}

static inline int config_read_file(config_t *config, const char *path) {
    FILE *f = fopen(path, "r");
    if (f == NULL) {
        snprintf(config->error, sizeof(config->error), "file I/O error");
        return CONFIG_FALSE;
    }
    char line[4096];
    int number = 0;
    while (fgets(line, sizeof(line), f) != NULL) {
        number++;
        char *p = line;
        while (isspace((unsigned char) *p)) p++;
        if (*p == '\0' || *p == '#' || (p[0] == '/' && p[1] == '/')) {
            continue;
        }
        char *eq = strpbrk(p, "=:");
        if (eq == NULL || eq == p) {
            snprintf(config->error, sizeof(config->error), "syntax error at line %d", number);
            fclose(f);
            return CONFIG_FALSE;
        }
        char *value = eq + 1;
        while (isspace((unsigned char) *value)) value++;
        char *end;
        int is_string = 0;
        if (*value == '"') {
            is_string = 1;
            value++;
            end = strchr(value, '"');
            if (end == NULL) {
                snprintf(config->error, sizeof(config->error), "syntax error at line %d", number);
                fclose(f);
                return CONFIG_FALSE;
            }
        } else {
            end = value + strcspn(value, ";#\r\n");
            if (end == value) {
                snprintf(config->error, sizeof(config->error), "syntax error at line %d", number);
                fclose(f);
                return CONFIG_FALSE;
            }
        }
        config_setting_t *grown = realloc(config->settings, (config->count + 1) * sizeof(config_setting_t));
        if (grown == NULL) {
            fclose(f);
            return CONFIG_FALSE;
        }
        config->settings = grown;
        config->settings[config->count].name = sdrm_cfg_copy(p, eq);
        config->settings[config->count].text = is_string ? sdrm_cfg_copy(value, end) : sdrm_cfg_copy(value, end);
        config->settings[config->count].is_string = is_string;
        config->count++;
    }
    fclose(f);
    return CONFIG_TRUE;
}

static inline config_setting_t *config_lookup(const config_t *config, const char *path) {
    for (size_t i = 0; i < config->count; i++) {
        if (strcmp(config->settings[i].name, path) == 0) {
            return &config->settings[i];
        }
    }
    return NULL;
}

static inline const char *config_setting_get_string(const config_setting_t *setting) { return setting->is_string ? setting->text : NULL; }

static inline int config_setting_get_int(const config_setting_t *setting) { return setting->is_string ? 0 : (int) strtol(setting->text, NULL, 0); }

static inline double config_setting_get_float(const config_setting_t *setting) { return setting->is_string ? 0.0 : strtod(setting->text, NULL); }

#endif
