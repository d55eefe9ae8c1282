/* fakex11.c — a display-less stand-in for libX11.so.6 / libXext.so.6, just enough for Mesa's xlib GLX state tracker to create a
 * context whose rendering goes to framebuffer objects.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/README.md).  This image ships a complete Mesa software rasteriser — llvmpipe, the
 * rasteriser underneath lavapipe — inside Nsight Compute (host/…/Mesa/libGL.so.1, Mesa 18.1.9, its xlib build), but neither libX11
 * nor an X server.  That libGL imports 30 Xlib symbols; they are all here.  No pixel ever reaches X: the caller renders into FBOs
 * and reads them back with glReadPixels, so the drawing entry points (XPutImage, XFillRectangle, …) are no-ops.
 *
 * The structures below are the public Xlib ABI (Xlib.h / Xutil.h / Xlibint.h of libX11 1.6, x86-64), restated from the ABI; the two
 * private offsets Mesa's glx_api.c reads directly (Display::ext_procs at 0x140, _XExtension::close_display / ::name at 0x48 / 0x60)
 * were checked against the disassembly of glXChooseVisual in that libGL and are static_asserted.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long XID, VisualID, Window, Drawable, Pixmap, Colormap, Font, Atom;
typedef char* XPointer;
typedef int Bool, Status;
typedef struct _XGC* GC;
struct _XDisplay;
typedef struct _XDisplay Display;

typedef struct { void* ext_data; VisualID visualid; int c_class; unsigned long red_mask, green_mask, blue_mask; int bits_per_rgb; int map_entries; } Visual;
typedef struct { int depth; int nvisuals; Visual* visuals; } Depth;
typedef struct {
	void* ext_data; Display* display; Window root; int width, height; int mwidth, mheight; int ndepths; Depth* depths; int root_depth;
	Visual* root_visual; GC default_gc; Colormap cmap; unsigned long white_pixel, black_pixel; int max_maps, min_maps; int backing_store;
	Bool save_unders; long root_input_mask;
} Screen;
typedef struct { void* ext_data; int depth; int bits_per_pixel; int scanline_pad; } ScreenFormat;
typedef struct { int extension, major_opcode, first_event, first_error; } XExtCodes;
typedef struct _XExten {
	struct _XExten* next; XExtCodes codes; void *create_GC, *copy_GC, *flush_GC, *free_GC, *create_Font, *free_Font;
	int (*close_display)(Display*, XExtCodes*); void *error, *error_string; char* name; void* error_values; void* before_flush; struct _XExten* next_flush;
} _XExtension;

struct _XDisplay {
	void* ext_data; void* free_funcs; int fd; int conn_checker; int proto_major_version, proto_minor_version; char* vendor;
	XID resource_base, resource_mask, resource_id; int resource_shift; XID (*resource_alloc)(Display*);
	int byte_order, bitmap_unit, bitmap_pad, bitmap_bit_order; int nformats; ScreenFormat* pixmap_format; int vnumber; int release;
	void *head, *tail; int qlen; unsigned long last_request_read, request; char *last_req, *buffer, *bufptr, *bufmax;
	unsigned max_request_size; void* db; int (*synchandler)(Display*); char* display_name; int default_screen; int nscreens;
	Screen* screens; unsigned long motion_buffer; volatile unsigned long flags; int min_keycode, max_keycode; void* keysyms;
	void* modifiermap; int keysyms_per_keycode; char* xdefaults; char* scratch_buffer; unsigned long scratch_length; int ext_number;
	_XExtension* ext_procs;
	char tail_pad[2048]; /* the rest of Xlibint's struct: never read by Mesa, zero */
};
_Static_assert(offsetof(struct _XDisplay, screens) == 232, "Xlib.h _XPrivDisplay::screens");
_Static_assert(offsetof(struct _XDisplay, ext_procs) == 0x140, "Xlibint.h _XDisplay::ext_procs (glXChooseVisual reads it at 0x140)");
_Static_assert(offsetof(_XExtension, close_display) == 0x48 && offsetof(_XExtension, name) == 0x60, "Xlibint.h _XExtension");
_Static_assert(sizeof(Screen) == 128, "Xlib.h Screen");

typedef struct { Visual* visual; VisualID visualid; int screen; int depth; int c_class; unsigned long red_mask, green_mask, blue_mask; int colormap_size; int bits_per_rgb; } XVisualInfo;
typedef struct _XImage {
	int width, height; int xoffset; int format; char* data; int byte_order; int bitmap_unit; int bitmap_bit_order; int bitmap_pad; int depth;
	int bytes_per_line; int bits_per_pixel; unsigned long red_mask, green_mask, blue_mask; XPointer obdata;
	struct funcs {
		struct _XImage* (*create_image)(void); int (*destroy_image)(struct _XImage*); unsigned long (*get_pixel)(struct _XImage*, int, int);
		int (*put_pixel)(struct _XImage*, int, int, unsigned long); struct _XImage* (*sub_image)(struct _XImage*, int, int, unsigned, unsigned);
		int (*add_pixel)(struct _XImage*, long);
	} f;
} XImage;
typedef struct {
	int x, y; int width, height; int border_width; int depth; Visual* visual; Window root; int c_class; int bit_gravity; int win_gravity;
	int backing_store; unsigned long backing_planes; unsigned long backing_pixel; Bool save_under; Colormap colormap; Bool map_installed;
	int map_state; long all_event_masks; long your_event_mask; long do_not_propagate_mask; Bool override_redirect; Screen* screen;
} XWindowAttributes;

enum { TrueColor = 4, ZPixmap = 2, LSBFirst = 0, kWindow = 0x400001, kColormap = 0x400002, kSize = 64 };

/* ---- one display, one screen, one 24-bit TrueColor visual (BGRX in memory, what every PC X server offers) */
static Visual g_visual = {0, 0x21, TrueColor, 0xff0000, 0x00ff00, 0x0000ff, 8, 256};
static Depth g_depth = {24, 1, &g_visual};
static Screen g_screen;
static ScreenFormat g_format = {0, 24, 32, 32};
static struct _XDisplay g_display;

Display* fakex_open_display(void) {
	memset(&g_display, 0, sizeof(g_display));
	memset(&g_screen, 0, sizeof(g_screen));
	g_screen.display = &g_display; g_screen.root = 0x100; g_screen.width = 1024; g_screen.height = 768; g_screen.mwidth = 270; g_screen.mheight = 203;
	g_screen.ndepths = 1; g_screen.depths = &g_depth; g_screen.root_depth = 24; g_screen.root_visual = &g_visual; g_screen.cmap = kColormap;
	g_screen.white_pixel = 0xffffff; g_screen.max_maps = 1; g_screen.min_maps = 1;
	g_display.proto_major_version = 11; g_display.vendor = (char*)"fakex11 (no server)"; g_display.byte_order = LSBFirst; g_display.bitmap_unit = 32;
	g_display.bitmap_pad = 32; g_display.bitmap_bit_order = LSBFirst; g_display.nformats = 1; g_display.pixmap_format = &g_format; g_display.release = 1;
	g_display.display_name = (char*)":fake"; g_display.nscreens = 1; g_display.screens = &g_screen; g_display.fd = -1;
	return &g_display;
}
Window fakex_window(void) { return kWindow; }

/* ---- what Mesa calls with a result it uses */
XVisualInfo* XGetVisualInfo(Display* d, long mask, XVisualInfo* tmpl, int* n) {
	(void)d;
	enum { IdMask = 1, ScreenMask = 2, DepthMask = 4, ClassMask = 8 };
	*n = 0;
	if ((mask & IdMask) && tmpl->visualid != g_visual.visualid) return NULL;
	if ((mask & ScreenMask) && tmpl->screen != 0) return NULL;
	if ((mask & DepthMask) && tmpl->depth != 24) return NULL;
	if ((mask & ClassMask) && tmpl->c_class != TrueColor) return NULL;
	XVisualInfo* v = (XVisualInfo*)calloc(1, sizeof(XVisualInfo));
	v->visual = &g_visual; v->visualid = g_visual.visualid; v->screen = 0; v->depth = 24; v->c_class = TrueColor;
	v->red_mask = g_visual.red_mask; v->green_mask = g_visual.green_mask; v->blue_mask = g_visual.blue_mask; v->colormap_size = 256; v->bits_per_rgb = 8;
	*n = 1;
	return v;
}
int XFree(void* p) { free(p); return 1; }

static int destroy_image(XImage* im) { if (im->data) free(im->data); free(im); return 1; }
static unsigned long get_pixel(XImage* im, int x, int y) { return im->data ? *(uint32_t*)(im->data + (size_t)y * im->bytes_per_line + (size_t)x * 4) : 0; }
static int put_pixel(XImage* im, int x, int y, unsigned long p) { if (im->data) *(uint32_t*)(im->data + (size_t)y * im->bytes_per_line + (size_t)x * 4) = (uint32_t)p; return 1; }
XImage* XCreateImage(Display* d, Visual* v, unsigned depth, int format, int offset, char* data, unsigned w, unsigned h, int pad, int bpl) {
	(void)d; (void)pad;
	XImage* im = (XImage*)calloc(1, sizeof(XImage));
	im->width = (int)w; im->height = (int)h; im->xoffset = offset; im->format = format; im->data = data; im->byte_order = LSBFirst; im->bitmap_unit = 32;
	im->bitmap_bit_order = LSBFirst; im->bitmap_pad = 32; im->depth = (int)depth; im->bits_per_pixel = depth > 8 ? 32 : 8;
	im->bytes_per_line = bpl ? bpl : (int)(w * (unsigned)im->bits_per_pixel / 8);
	if (v) { im->red_mask = v->red_mask; im->green_mask = v->green_mask; im->blue_mask = v->blue_mask; }
	im->f.destroy_image = destroy_image; im->f.get_pixel = get_pixel; im->f.put_pixel = put_pixel;
	return im;
}
XImage* XShmCreateImage(Display* d, Visual* v, unsigned depth, int format, char* data, void* shminfo, unsigned w, unsigned h) {
	(void)shminfo;
	return XCreateImage(d, v, depth, format, 0, data, w, h, 32, 0);
}
XImage* XGetImage(Display* d, Drawable dr, int x, int y, unsigned w, unsigned h, unsigned long planes, int format) {
	(void)dr; (void)x; (void)y; (void)planes;
	return XCreateImage(d, &g_visual, 24, format, 0, (char*)calloc((size_t)w * h, 4), w, h, 32, 0);
}

Status XGetGeometry(Display* d, Drawable dr, Window* root, int* x, int* y, unsigned* w, unsigned* h, unsigned* border, unsigned* depth) {
	(void)d; (void)dr;
	if (root) *root = g_screen.root;
	if (x) *x = 0;
	if (y) *y = 0;
	if (w) *w = kSize;
	if (h) *h = kSize;
	if (border) *border = 0;
	if (depth) *depth = 24;
	return 1;
}
Status XGetWindowAttributes(Display* d, Window w, XWindowAttributes* a) {
	(void)d; (void)w;
	memset(a, 0, sizeof(*a));
	a->width = kSize; a->height = kSize; a->depth = 24; a->visual = &g_visual; a->root = g_screen.root; a->c_class = 1 /* InputOutput */;
	a->colormap = kColormap; a->map_installed = 1; a->map_state = 2 /* IsViewable */; a->screen = &g_screen;
	return 1;
}

XExtCodes* XAddExtension(Display* d) {
	_XExtension* e = (_XExtension*)calloc(1, sizeof(_XExtension));
	e->codes.extension = d->ext_number++;
	e->next = d->ext_procs;
	d->ext_procs = e;
	return &e->codes;
}
Bool XQueryExtension(Display* d, const char* name, int* op, int* ev, int* err) { (void)d; (void)name; (void)op; (void)ev; (void)err; return 0; } /* no MIT-SHM, no GLX */

/* ---- resources that are only handles here */
GC XCreateGC(Display* d, Drawable dr, unsigned long mask, void* values) { (void)d; (void)dr; (void)mask; (void)values; return (GC)calloc(1, 128); }
int XFreeGC(Display* d, GC gc) { (void)d; free(gc); return 1; }
Colormap XCreateColormap(Display* d, Window w, Visual* v, int alloc) { (void)d; (void)w; (void)v; (void)alloc; return kColormap; }
Pixmap XCreatePixmap(Display* d, Drawable dr, unsigned w, unsigned h, unsigned depth) { (void)d; (void)dr; (void)w; (void)h; (void)depth; return 0x400010; }
int XFreePixmap(Display* d, Pixmap p) { (void)d; (void)p; return 1; }

/* ---- no-ops (drawing to the display, fonts, synchronisation with a server that is not there) */
typedef int (*XErrorHandler)(Display*, void*);
XErrorHandler XSetErrorHandler(XErrorHandler h) { (void)h; return NULL; }
int XSync(Display* d, Bool discard) { (void)d; (void)discard; return 1; }
int XFlush(Display* d) { (void)d; return 1; }
int (*XSynchronize(Display* d, Bool onoff))(Display*) { (void)d; (void)onoff; return NULL; }
int XSetForeground(Display* d, GC gc, unsigned long fg) { (void)d; (void)gc; (void)fg; return 1; }
int XSetFunction(Display* d, GC gc, int fn) { (void)d; (void)gc; (void)fn; return 1; }
int XFillRectangle(Display* d, Drawable dr, GC gc, int x, int y, unsigned w, unsigned h) { (void)d; (void)dr; (void)gc; (void)x; (void)y; (void)w; (void)h; return 1; }
int XPutImage(Display* d, Drawable dr, GC gc, XImage* im, int sx, int sy, int dx, int dy, unsigned w, unsigned h) { (void)d; (void)dr; (void)gc; (void)im; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h; return 1; }
Bool XShmPutImage(Display* d, Drawable dr, GC gc, XImage* im, int sx, int sy, int dx, int dy, unsigned w, unsigned h, Bool ev) { (void)ev; return XPutImage(d, dr, gc, im, sx, sy, dx, dy, w, h); }
Bool XShmAttach(Display* d, void* info) { (void)d; (void)info; return 0; }
int XDrawString16(Display* d, Drawable dr, GC gc, int x, int y, const void* s, int n) { (void)d; (void)dr; (void)gc; (void)x; (void)y; (void)s; (void)n; return 1; }
void* XQueryFont(Display* d, XID id) { (void)d; (void)id; return NULL; }
int XFreeFontInfo(char** names, void* info, int n) { (void)names; (void)info; (void)n; return 1; }

/* Xlib's optional thread lock hooks: unset = single-threaded Xlib, which is what the _XLockMutex() macro checks for */
void (*_XLockMutex_fn)(void*) = NULL;
void (*_XUnlockMutex_fn)(void*) = NULL;
void* _Xglobal_lock = NULL;
