/*
 * aidax_lv2_abi.h -- layout-exact restatement of the part of the public LV2 1.18 C ABI
 * (ISC-licensed specification, http://lv2plug.in) that the rt-neural-generic plugin
 * surface touches.  The LV2 development headers are not installed in the build image and
 * there is no network, so the build carries its own declarations.  Everything a real LV2
 * host passes across the plugin boundary (LV2_Descriptor, LV2_Feature, URID map, atoms,
 * worker, state, log) has the exact field order, types and enum values of the spec, so a
 * stock host (jalv, mod-host, Ardour) can load the .so built against this header.
 *
 * Replaces, for this repository only, the reference's includes at
 *   rt-neural-generic/src/rt-neural-generic.h:33-41 and rt-neural-generic/src/uris.h:21-24
 * (<lv2/atom/forge.h> <lv2/atom/util.h> <lv2/log/log.h> <lv2/log/logger.h> <lv2/midi/midi.h>
 *  <lv2/patch/patch.h> <lv2/state/state.h> <lv2/urid/urid.h> <lv2/worker/worker.h>
 *  <lv2/parameters/parameters.h>).
 *
 * The atom forge and the logger helper are header-only conveniences in LV2 (never passed
 * to the host by layout), so they are re-implemented here with the same call signatures
 * and the same byte output, not the same private struct layout.
 */
#ifndef AIDAX_LV2_ABI_H
#define AIDAX_LV2_ABI_H

#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifdef __cplusplus
#define LV2_SYMBOL_EXTERN extern "C"
#else
#define LV2_SYMBOL_EXTERN
#endif
#define LV2_SYMBOL_EXPORT LV2_SYMBOL_EXTERN __attribute__((visibility("default")))

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ core */
#define LV2_CORE_URI "http://lv2plug.in/ns/lv2core"
#define LV2_CORE_PREFIX LV2_CORE_URI "#"

typedef void* LV2_Handle;

typedef struct LV2_Feature {
    const char* URI;
    void* data;
} LV2_Feature;

typedef struct LV2_Descriptor {
    const char* URI;
    LV2_Handle (*instantiate)(const struct LV2_Descriptor* descriptor, double sample_rate,
                              const char* bundle_path, const LV2_Feature* const* features);
    void (*connect_port)(LV2_Handle instance, uint32_t port, void* data_location);
    void (*activate)(LV2_Handle instance);
    void (*run)(LV2_Handle instance, uint32_t sample_count);
    void (*deactivate)(LV2_Handle instance);
    void (*cleanup)(LV2_Handle instance);
    const void* (*extension_data)(const char* uri);
} LV2_Descriptor;

typedef const LV2_Descriptor* (*LV2_Descriptor_Function)(uint32_t index);

/* ------------------------------------------------------------------ urid */
#define LV2_URID_URI "http://lv2plug.in/ns/ext/urid"
#define LV2_URID_PREFIX LV2_URID_URI "#"
#define LV2_URID__map LV2_URID_PREFIX "map"
#define LV2_URID__unmap LV2_URID_PREFIX "unmap"

typedef void* LV2_URID_Map_Handle;
typedef void* LV2_URID_Unmap_Handle;
typedef uint32_t LV2_URID;

typedef struct LV2_URID_Map {
    LV2_URID_Map_Handle handle;
    LV2_URID (*map)(LV2_URID_Map_Handle handle, const char* uri);
} LV2_URID_Map;

typedef struct LV2_URID_Unmap {
    LV2_URID_Unmap_Handle handle;
    const char* (*unmap)(LV2_URID_Unmap_Handle handle, LV2_URID urid);
} LV2_URID_Unmap;

/* ------------------------------------------------------------------ atom */
#define LV2_ATOM_URI "http://lv2plug.in/ns/ext/atom"
#define LV2_ATOM_PREFIX LV2_ATOM_URI "#"
#define LV2_ATOM__Atom LV2_ATOM_PREFIX "Atom"
#define LV2_ATOM__AtomPort LV2_ATOM_PREFIX "AtomPort"
#define LV2_ATOM__Blank LV2_ATOM_PREFIX "Blank"
#define LV2_ATOM__Bool LV2_ATOM_PREFIX "Bool"
#define LV2_ATOM__Chunk LV2_ATOM_PREFIX "Chunk"
#define LV2_ATOM__Double LV2_ATOM_PREFIX "Double"
#define LV2_ATOM__Event LV2_ATOM_PREFIX "Event"
#define LV2_ATOM__Float LV2_ATOM_PREFIX "Float"
#define LV2_ATOM__Int LV2_ATOM_PREFIX "Int"
#define LV2_ATOM__Literal LV2_ATOM_PREFIX "Literal"
#define LV2_ATOM__Long LV2_ATOM_PREFIX "Long"
#define LV2_ATOM__Number LV2_ATOM_PREFIX "Number"
#define LV2_ATOM__Object LV2_ATOM_PREFIX "Object"
#define LV2_ATOM__Path LV2_ATOM_PREFIX "Path"
#define LV2_ATOM__Property LV2_ATOM_PREFIX "Property"
#define LV2_ATOM__Resource LV2_ATOM_PREFIX "Resource"
#define LV2_ATOM__Sequence LV2_ATOM_PREFIX "Sequence"
#define LV2_ATOM__Sound LV2_ATOM_PREFIX "Sound"
#define LV2_ATOM__String LV2_ATOM_PREFIX "String"
#define LV2_ATOM__Tuple LV2_ATOM_PREFIX "Tuple"
#define LV2_ATOM__URI LV2_ATOM_PREFIX "URI"
#define LV2_ATOM__URID LV2_ATOM_PREFIX "URID"
#define LV2_ATOM__Vector LV2_ATOM_PREFIX "Vector"
#define LV2_ATOM__atomTransfer LV2_ATOM_PREFIX "atomTransfer"
#define LV2_ATOM__beatTime LV2_ATOM_PREFIX "beatTime"
#define LV2_ATOM__bufferType LV2_ATOM_PREFIX "bufferType"
#define LV2_ATOM__childType LV2_ATOM_PREFIX "childType"
#define LV2_ATOM__eventTransfer LV2_ATOM_PREFIX "eventTransfer"
#define LV2_ATOM__frameTime LV2_ATOM_PREFIX "frameTime"
#define LV2_ATOM__supports LV2_ATOM_PREFIX "supports"
#define LV2_ATOM__timeUnit LV2_ATOM_PREFIX "timeUnit"

#define LV2_ATOM_REFERENCE_TYPE 0

typedef struct LV2_Atom {
    uint32_t size; /* size of the body in bytes, not including this header */
    uint32_t type; /* URID */
} LV2_Atom;

typedef struct LV2_Atom_Int { LV2_Atom atom; int32_t body; } LV2_Atom_Int;
typedef struct LV2_Atom_Long { LV2_Atom atom; int64_t body; } LV2_Atom_Long;
typedef struct LV2_Atom_Float { LV2_Atom atom; float body; } LV2_Atom_Float;
typedef struct LV2_Atom_Double { LV2_Atom atom; double body; } LV2_Atom_Double;
typedef LV2_Atom_Int LV2_Atom_Bool;
typedef struct LV2_Atom_URID { LV2_Atom atom; LV2_URID body; } LV2_Atom_URID;
typedef struct LV2_Atom_String { LV2_Atom atom; /* chars follow */ } LV2_Atom_String;

typedef struct LV2_Atom_Property_Body {
    LV2_URID key;
    LV2_URID context;
    LV2_Atom value; /* value body follows */
} LV2_Atom_Property_Body;

typedef struct LV2_Atom_Object_Body {
    LV2_URID id;
    LV2_URID otype;
} LV2_Atom_Object_Body;

typedef struct LV2_Atom_Object {
    LV2_Atom atom;
    LV2_Atom_Object_Body body;
} LV2_Atom_Object;

typedef struct LV2_Atom_Event {
    union {
        int64_t frames;
        double beats;
    } time;
    LV2_Atom body; /* body bytes follow */
} LV2_Atom_Event;

typedef struct LV2_Atom_Sequence_Body {
    LV2_URID unit;
    uint32_t pad;
} LV2_Atom_Sequence_Body;

typedef struct LV2_Atom_Sequence {
    LV2_Atom atom;
    LV2_Atom_Sequence_Body body;
} LV2_Atom_Sequence;

/* ------------------------------------------------------------- atom/util */
static inline uint32_t lv2_atom_pad_size(uint32_t size) { return (size + 7U) & (~7U); }
static inline uint32_t lv2_atom_total_size(const LV2_Atom* atom) { return (uint32_t)sizeof(LV2_Atom) + atom->size; }
static inline bool lv2_atom_is_null(const LV2_Atom* atom) { return !atom || (atom->type == 0 && atom->size == 0); }

#define LV2_ATOM_CONTENTS(type, atom) ((void*)((uint8_t*)(atom) + sizeof(type)))
#define LV2_ATOM_CONTENTS_CONST(type, atom) ((const void*)((const uint8_t*)(atom) + sizeof(type)))
#define LV2_ATOM_BODY(atom) LV2_ATOM_CONTENTS(LV2_Atom, atom)
#define LV2_ATOM_BODY_CONST(atom) LV2_ATOM_CONTENTS_CONST(LV2_Atom, atom)

static inline LV2_Atom_Event* lv2_atom_sequence_begin(const LV2_Atom_Sequence_Body* body) {
    return (LV2_Atom_Event*)(body + 1);
}
static inline LV2_Atom_Event* lv2_atom_sequence_end(const LV2_Atom_Sequence_Body* body, uint32_t size) {
    return (LV2_Atom_Event*)((const uint8_t*)body + lv2_atom_pad_size(size));
}
static inline bool lv2_atom_sequence_is_end(const LV2_Atom_Sequence_Body* body, uint32_t size, const LV2_Atom_Event* i) {
    return (const uint8_t*)i >= ((const uint8_t*)body + size);
}
static inline LV2_Atom_Event* lv2_atom_sequence_next(const LV2_Atom_Event* i) {
    return (LV2_Atom_Event*)((const uint8_t*)i + sizeof(LV2_Atom_Event) + lv2_atom_pad_size(i->body.size));
}
#define LV2_ATOM_SEQUENCE_FOREACH(seq, iter)                                              \
    for (LV2_Atom_Event * iter = lv2_atom_sequence_begin(&(seq)->body);                   \
         !lv2_atom_sequence_is_end(&(seq)->body, (seq)->atom.size, (iter));               \
         (iter) = lv2_atom_sequence_next(iter))

static inline void lv2_atom_sequence_clear(LV2_Atom_Sequence* seq) { seq->atom.size = sizeof(LV2_Atom_Sequence_Body); }

static inline LV2_Atom_Property_Body* lv2_atom_object_begin(const LV2_Atom_Object_Body* body) {
    return (LV2_Atom_Property_Body*)(body + 1);
}
static inline bool lv2_atom_object_is_end(const LV2_Atom_Object_Body* body, uint32_t size, const LV2_Atom_Property_Body* i) {
    return (const uint8_t*)i >= ((const uint8_t*)body + size);
}
static inline LV2_Atom_Property_Body* lv2_atom_object_next(const LV2_Atom_Property_Body* i) {
    const LV2_Atom* const value = (const LV2_Atom*)((const uint8_t*)i + 2 * sizeof(uint32_t));
    return (LV2_Atom_Property_Body*)((const uint8_t*)i + lv2_atom_pad_size((uint32_t)sizeof(LV2_Atom_Property_Body) + value->size));
}
#define LV2_ATOM_OBJECT_FOREACH(obj, iter)                                                \
    for (LV2_Atom_Property_Body * iter = lv2_atom_object_begin(&(obj)->body);             \
         !lv2_atom_object_is_end(&(obj)->body, (obj)->atom.size, (iter));                 \
         (iter) = lv2_atom_object_next(iter))

/* Variable-argument object query: (key, const LV2_Atom** out)..., terminated by key 0.
 * Returns the number of matches. */
static inline int lv2_atom_object_get(const LV2_Atom_Object* object, ...) {
    int matches = 0;
    int n_queries = 0;
    va_list args;
    va_start(args, object);
    for (n_queries = 0; va_arg(args, uint32_t); ++n_queries) {
        if (!va_arg(args, const LV2_Atom**)) {
            va_end(args);
            return -1;
        }
    }
    va_end(args);

    LV2_ATOM_OBJECT_FOREACH(object, prop) {
        va_start(args, object);
        for (int i = 0; i < n_queries; ++i) {
            const uint32_t qkey = va_arg(args, uint32_t);
            const LV2_Atom** qval = va_arg(args, const LV2_Atom**);
            if (qkey == prop->key && !*qval) {
                *qval = &prop->value;
                if (++matches == n_queries) {
                    va_end(args);
                    return matches;
                }
                break;
            }
        }
        va_end(args);
    }
    return matches;
}

/* ------------------------------------------------------------ atom/forge */
typedef void* LV2_Atom_Forge_Sink_Handle;
typedef intptr_t LV2_Atom_Forge_Ref;

typedef struct LV2_Atom_Forge_Frame {
    struct LV2_Atom_Forge_Frame* parent;
    LV2_Atom_Forge_Ref ref; /* byte offset + 1 of the container atom inside the buffer; 0 = none */
} LV2_Atom_Forge_Frame;

typedef struct LV2_Atom_Forge {
    uint8_t* buf;
    uint32_t offset;
    uint32_t size;
    LV2_Atom_Forge_Frame* stack;
    LV2_URID Blank, Bool, Chunk, Double, Float, Int, Long, Literal, Object, Path, Property,
        Resource, Sequence, String, Tuple, URI, URID, Vector;
} LV2_Atom_Forge;

static inline void lv2_atom_forge_set_buffer(LV2_Atom_Forge* forge, uint8_t* buf, size_t size) {
    forge->buf = buf;
    forge->size = (uint32_t)size;
    forge->offset = 0;
    forge->stack = NULL;
}

static inline void lv2_atom_forge_init(LV2_Atom_Forge* forge, LV2_URID_Map* map) {
    lv2_atom_forge_set_buffer(forge, NULL, 0);
    forge->Blank = map->map(map->handle, LV2_ATOM__Blank);
    forge->Bool = map->map(map->handle, LV2_ATOM__Bool);
    forge->Chunk = map->map(map->handle, LV2_ATOM__Chunk);
    forge->Double = map->map(map->handle, LV2_ATOM__Double);
    forge->Float = map->map(map->handle, LV2_ATOM__Float);
    forge->Int = map->map(map->handle, LV2_ATOM__Int);
    forge->Long = map->map(map->handle, LV2_ATOM__Long);
    forge->Literal = map->map(map->handle, LV2_ATOM__Literal);
    forge->Object = map->map(map->handle, LV2_ATOM__Object);
    forge->Path = map->map(map->handle, LV2_ATOM__Path);
    forge->Property = map->map(map->handle, LV2_ATOM__Property);
    forge->Resource = map->map(map->handle, LV2_ATOM__Resource);
    forge->Sequence = map->map(map->handle, LV2_ATOM__Sequence);
    forge->String = map->map(map->handle, LV2_ATOM__String);
    forge->Tuple = map->map(map->handle, LV2_ATOM__Tuple);
    forge->URI = map->map(map->handle, LV2_ATOM__URI);
    forge->URID = map->map(map->handle, LV2_ATOM__URID);
    forge->Vector = map->map(map->handle, LV2_ATOM__Vector);
}

static inline LV2_Atom* lv2_atom_forge_deref(LV2_Atom_Forge* forge, LV2_Atom_Forge_Ref ref) {
    return (LV2_Atom*)(forge->buf + (ref - 1));
}

static inline bool lv2_atom_forge_is_object_type(const LV2_Atom_Forge* forge, uint32_t type) {
    return type == forge->Object || type == forge->Blank || type == forge->Resource;
}

static inline LV2_Atom_Forge_Ref lv2_atom_forge_push(LV2_Atom_Forge* forge, LV2_Atom_Forge_Frame* frame, LV2_Atom_Forge_Ref ref) {
    frame->parent = forge->stack;
    frame->ref = ref;
    if (ref) forge->stack = frame;
    return ref;
}

static inline void lv2_atom_forge_pop(LV2_Atom_Forge* forge, LV2_Atom_Forge_Frame* frame) {
    if (frame->ref) forge->stack = frame->parent;
}

/* Append raw bytes; every open container on the frame stack grows by `size`. */
static inline LV2_Atom_Forge_Ref lv2_atom_forge_raw(LV2_Atom_Forge* forge, const void* data, uint32_t size) {
    if (!forge->buf || forge->offset + size > forge->size) return 0;
    uint8_t* mem = forge->buf + forge->offset;
    const LV2_Atom_Forge_Ref out = (LV2_Atom_Forge_Ref)forge->offset + 1;
    forge->offset += size;
    if (data) memcpy(mem, data, size); else memset(mem, 0, size);
    for (LV2_Atom_Forge_Frame* f = forge->stack; f; f = f->parent)
        lv2_atom_forge_deref(forge, f->ref)->size += size;
    return out;
}

static inline void lv2_atom_forge_pad(LV2_Atom_Forge* forge, uint32_t written) {
    const uint64_t pad = 0;
    const uint32_t pad_size = lv2_atom_pad_size(written) - written;
    lv2_atom_forge_raw(forge, &pad, pad_size);
}

static inline LV2_Atom_Forge_Ref lv2_atom_forge_write(LV2_Atom_Forge* forge, const void* data, uint32_t size) {
    const LV2_Atom_Forge_Ref out = lv2_atom_forge_raw(forge, data, size);
    if (out) lv2_atom_forge_pad(forge, size);
    return out;
}

static inline LV2_Atom_Forge_Ref lv2_atom_forge_urid(LV2_Atom_Forge* forge, LV2_URID id) {
    const LV2_Atom_URID a = { { (uint32_t)sizeof(id), forge->URID }, id };
    return lv2_atom_forge_write(forge, &a, (uint32_t)sizeof(a));
}

static inline LV2_Atom_Forge_Ref lv2_atom_forge_float(LV2_Atom_Forge* forge, float val) {
    const LV2_Atom_Float a = { { (uint32_t)sizeof(val), forge->Float }, val };
    return lv2_atom_forge_write(forge, &a, (uint32_t)sizeof(a));
}

/* String-like atoms: header with size len+1, body, NUL, pad to 8. */
static inline LV2_Atom_Forge_Ref lv2_atom_forge_typed_string(LV2_Atom_Forge* forge, uint32_t type, const char* str, uint32_t len) {
    const LV2_Atom_String a = { { len + 1, type } };
    LV2_Atom_Forge_Ref out = lv2_atom_forge_raw(forge, &a, (uint32_t)sizeof(a));
    if (out) {
        if (!lv2_atom_forge_raw(forge, str, len) || !lv2_atom_forge_raw(forge, "", 1)) {
            LV2_Atom* atom = lv2_atom_forge_deref(forge, out);
            atom->size = atom->type = 0;
            out = 0;
        } else {
            lv2_atom_forge_pad(forge, (uint32_t)sizeof(a) + len + 1);
        }
    }
    return out;
}
static inline LV2_Atom_Forge_Ref lv2_atom_forge_path(LV2_Atom_Forge* forge, const char* str, uint32_t len) {
    return lv2_atom_forge_typed_string(forge, forge->Path, str, len);
}
static inline LV2_Atom_Forge_Ref lv2_atom_forge_string(LV2_Atom_Forge* forge, const char* str, uint32_t len) {
    return lv2_atom_forge_typed_string(forge, forge->String, str, len);
}

static inline LV2_Atom_Forge_Ref lv2_atom_forge_sequence_head(LV2_Atom_Forge* forge, LV2_Atom_Forge_Frame* frame, uint32_t unit) {
    const LV2_Atom_Sequence a = { { (uint32_t)sizeof(LV2_Atom_Sequence_Body), forge->Sequence }, { unit, 0 } };
    return lv2_atom_forge_push(forge, frame, lv2_atom_forge_write(forge, &a, (uint32_t)sizeof(a)));
}

static inline LV2_Atom_Forge_Ref lv2_atom_forge_object(LV2_Atom_Forge* forge, LV2_Atom_Forge_Frame* frame, LV2_URID id, LV2_URID otype) {
    const LV2_Atom_Object a = { { (uint32_t)sizeof(LV2_Atom_Object_Body), forge->Object }, { id, otype } };
    return lv2_atom_forge_push(forge, frame, lv2_atom_forge_write(forge, &a, (uint32_t)sizeof(a)));
}

static inline LV2_Atom_Forge_Ref lv2_atom_forge_key(LV2_Atom_Forge* forge, LV2_URID key) {
    const uint32_t body[2] = { key, 0 }; /* key, context; the value atom follows */
    return lv2_atom_forge_write(forge, body, (uint32_t)sizeof(body));
}

static inline LV2_Atom_Forge_Ref lv2_atom_forge_frame_time(LV2_Atom_Forge* forge, int64_t frames) {
    return lv2_atom_forge_write(forge, &frames, (uint32_t)sizeof(frames));
}

/* ------------------------------------------------------------------- log */
#define LV2_LOG_URI "http://lv2plug.in/ns/ext/log"
#define LV2_LOG_PREFIX LV2_LOG_URI "#"
#define LV2_LOG__Entry LV2_LOG_PREFIX "Entry"
#define LV2_LOG__Error LV2_LOG_PREFIX "Error"
#define LV2_LOG__Note LV2_LOG_PREFIX "Note"
#define LV2_LOG__Trace LV2_LOG_PREFIX "Trace"
#define LV2_LOG__Warning LV2_LOG_PREFIX "Warning"
#define LV2_LOG__log LV2_LOG_PREFIX "log"

typedef void* LV2_Log_Handle;

typedef struct LV2_Log_Log {
    LV2_Log_Handle handle;
    int (*printf)(LV2_Log_Handle handle, LV2_URID type, const char* fmt, ...);
    int (*vprintf)(LV2_Log_Handle handle, LV2_URID type, const char* fmt, va_list ap);
} LV2_Log_Log;

typedef struct LV2_Log_Logger {
    LV2_Log_Log* log;
    LV2_URID Error, Note, Trace, Warning;
} LV2_Log_Logger;

static inline void lv2_log_logger_set_map(LV2_Log_Logger* logger, LV2_URID_Map* map) {
    if (map) {
        logger->Error = map->map(map->handle, LV2_LOG__Error);
        logger->Note = map->map(map->handle, LV2_LOG__Note);
        logger->Trace = map->map(map->handle, LV2_LOG__Trace);
        logger->Warning = map->map(map->handle, LV2_LOG__Warning);
    } else {
        logger->Error = logger->Note = logger->Trace = logger->Warning = 0;
    }
}
static inline void lv2_log_logger_init(LV2_Log_Logger* logger, LV2_URID_Map* map, LV2_Log_Log* log) {
    logger->log = log;
    lv2_log_logger_set_map(logger, map);
}
static inline int lv2_log_vprintf(LV2_Log_Logger* logger, LV2_URID type, const char* fmt, va_list args) {
    return (logger && logger->log) ? logger->log->vprintf(logger->log->handle, type, fmt, args)
                                   : vfprintf(stderr, fmt, args);
}
#define AIDAX_LV2_LOG_FN(name, field)                                                    \
    __attribute__((format(printf, 2, 3))) static inline int name(LV2_Log_Logger* logger, const char* fmt, ...) { \
        va_list args;                                                                    \
        va_start(args, fmt);                                                             \
        const int ret = lv2_log_vprintf(logger, logger->field, fmt, args);               \
        va_end(args);                                                                    \
        return ret;                                                                      \
    }
AIDAX_LV2_LOG_FN(lv2_log_error, Error)
AIDAX_LV2_LOG_FN(lv2_log_note, Note)
AIDAX_LV2_LOG_FN(lv2_log_trace, Trace)
AIDAX_LV2_LOG_FN(lv2_log_warning, Warning)
#undef AIDAX_LV2_LOG_FN

/* ---------------------------------------------------------------- worker */
#define LV2_WORKER_URI "http://lv2plug.in/ns/ext/worker"
#define LV2_WORKER_PREFIX LV2_WORKER_URI "#"
#define LV2_WORKER__interface LV2_WORKER_PREFIX "interface"
#define LV2_WORKER__schedule LV2_WORKER_PREFIX "schedule"

typedef enum {
    LV2_WORKER_SUCCESS = 0,
    LV2_WORKER_ERR_UNKNOWN = 1,
    LV2_WORKER_ERR_NO_SPACE = 2
} LV2_Worker_Status;

typedef void* LV2_Worker_Respond_Handle;
typedef LV2_Worker_Status (*LV2_Worker_Respond_Function)(LV2_Worker_Respond_Handle handle, uint32_t size, const void* data);

typedef struct LV2_Worker_Interface {
    LV2_Worker_Status (*work)(LV2_Handle instance, LV2_Worker_Respond_Function respond,
                              LV2_Worker_Respond_Handle handle, uint32_t size, const void* data);
    LV2_Worker_Status (*work_response)(LV2_Handle instance, uint32_t size, const void* body);
    LV2_Worker_Status (*end_run)(LV2_Handle instance);
} LV2_Worker_Interface;

typedef void* LV2_Worker_Schedule_Handle;

typedef struct LV2_Worker_Schedule {
    LV2_Worker_Schedule_Handle handle;
    LV2_Worker_Status (*schedule_work)(LV2_Worker_Schedule_Handle handle, uint32_t size, const void* data);
} LV2_Worker_Schedule;

/* ----------------------------------------------------------------- state */
#define LV2_STATE_URI "http://lv2plug.in/ns/ext/state"
#define LV2_STATE_PREFIX LV2_STATE_URI "#"
#define LV2_STATE__State LV2_STATE_PREFIX "State"
#define LV2_STATE__interface LV2_STATE_PREFIX "interface"
#define LV2_STATE__loadDefaultState LV2_STATE_PREFIX "loadDefaultState"
#define LV2_STATE__freePath LV2_STATE_PREFIX "freePath"
#define LV2_STATE__makePath LV2_STATE_PREFIX "makePath"
#define LV2_STATE__mapPath LV2_STATE_PREFIX "mapPath"
#define LV2_STATE__state LV2_STATE_PREFIX "state"
#define LV2_STATE__threadSafeRestore LV2_STATE_PREFIX "threadSafeRestore"

typedef void* LV2_State_Handle;
typedef void* LV2_State_Free_Path_Handle;
typedef void* LV2_State_Map_Path_Handle;
typedef void* LV2_State_Make_Path_Handle;

typedef enum {
    LV2_STATE_IS_POD = 1U << 0U,
    LV2_STATE_IS_PORTABLE = 1U << 1U,
    LV2_STATE_IS_NATIVE = 1U << 2U
} LV2_State_Flags;

typedef enum {
    LV2_STATE_SUCCESS = 0,
    LV2_STATE_ERR_UNKNOWN = 1,
    LV2_STATE_ERR_BAD_TYPE = 2,
    LV2_STATE_ERR_BAD_FLAGS = 3,
    LV2_STATE_ERR_NO_FEATURE = 4,
    LV2_STATE_ERR_NO_PROPERTY = 5,
    LV2_STATE_ERR_NO_SPACE = 6
} LV2_State_Status;

typedef LV2_State_Status (*LV2_State_Store_Function)(LV2_State_Handle handle, uint32_t key, const void* value,
                                                     size_t size, uint32_t type, uint32_t flags);
typedef const void* (*LV2_State_Retrieve_Function)(LV2_State_Handle handle, uint32_t key, size_t* size,
                                                   uint32_t* type, uint32_t* flags);

typedef struct LV2_State_Interface {
    LV2_State_Status (*save)(LV2_Handle instance, LV2_State_Store_Function store, LV2_State_Handle handle,
                             uint32_t flags, const LV2_Feature* const* features);
    LV2_State_Status (*restore)(LV2_Handle instance, LV2_State_Retrieve_Function retrieve, LV2_State_Handle handle,
                                uint32_t flags, const LV2_Feature* const* features);
} LV2_State_Interface;

typedef struct LV2_State_Map_Path {
    LV2_State_Map_Path_Handle handle;
    char* (*abstract_path)(LV2_State_Map_Path_Handle handle, const char* absolute_path);
    char* (*absolute_path)(LV2_State_Map_Path_Handle handle, const char* abstract_path);
} LV2_State_Map_Path;

typedef struct LV2_State_Make_Path {
    LV2_State_Make_Path_Handle handle;
    char* (*path)(LV2_State_Make_Path_Handle handle, const char* path);
} LV2_State_Make_Path;

typedef struct LV2_State_Free_Path {
    LV2_State_Free_Path_Handle handle;
    void (*free_path)(LV2_State_Free_Path_Handle handle, char* path);
} LV2_State_Free_Path;

/* ------------------------------------------------- patch / midi / params */
#define LV2_PATCH_URI "http://lv2plug.in/ns/ext/patch"
#define LV2_PATCH_PREFIX LV2_PATCH_URI "#"
#define LV2_PATCH__Get LV2_PATCH_PREFIX "Get"
#define LV2_PATCH__Set LV2_PATCH_PREFIX "Set"
#define LV2_PATCH__Message LV2_PATCH_PREFIX "Message"
#define LV2_PATCH__property LV2_PATCH_PREFIX "property"
#define LV2_PATCH__value LV2_PATCH_PREFIX "value"
#define LV2_PATCH__writable LV2_PATCH_PREFIX "writable"

#define LV2_MIDI_URI "http://lv2plug.in/ns/ext/midi"
#define LV2_MIDI_PREFIX LV2_MIDI_URI "#"
#define LV2_MIDI__MidiEvent LV2_MIDI_PREFIX "MidiEvent"

#define LV2_PARAMETERS_URI "http://lv2plug.in/ns/ext/parameters"
#define LV2_PARAMETERS_PREFIX LV2_PARAMETERS_URI "#"
#define LV2_PARAMETERS__gain LV2_PARAMETERS_PREFIX "gain"

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* AIDAX_LV2_ABI_H */
