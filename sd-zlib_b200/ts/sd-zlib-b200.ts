/*
 * sd-zlib-b200.ts - TypeScript facade with the public surface of @stardazed/zlib for the
 * inflate + checksum path (dist/sd-zlib.d.ts:11-148 of the reference), backed by the N-API shim
 * (sdz_napi.c) over libsdzcuda.so.  Deflater / deflate stay the reference's own CPU code and are
 * re-exported unchanged by the package entry.
 *
 * NOT BUILT OR RUN IN THIS REPOSITORY (no Node.js / tsc in the image).  The Python mirror
 * sd-zlib_b200/host/sdzlib implements exactly this logic and is what the tests exercise.
 */
import { createRequire } from "node:module";
const native = createRequire(import.meta.url)("./sdz_napi.node") as {
	adler32(buf: Uint8Array, seed: number): number;
	crc32(buf: Uint8Array, seed: number): number;
	inflateSizes(bufs: Uint8Array[], dicts: (Uint8Array | undefined)[], modes: Uint8Array): Float64Array;
	inflateBatch(bufs: Uint8Array[], dicts: (Uint8Array | undefined)[], modes: Uint8Array, out: Uint8Array,
		outOff: BigUint64Array, outCap: BigUint64Array, records: Uint8Array): number;
	inflaterNew(raw: boolean, dict: Uint8Array | undefined): unknown;
	inflaterAppend(handle: unknown, chunk: Uint8Array, record: Uint8Array): Uint8Array;
	inflaterFinish(handle: unknown, record: Uint8Array): Uint8Array;
};

export interface InflaterOptions { raw?: boolean; dictionary?: BufferSource; }
export interface InflateResult {
	success: boolean; complete: boolean;
	checksum: "match" | "mismatch" | "unchecked"; fileSize: "match" | "mismatch" | "unchecked";
	fileName: string; modDate: Date | undefined;
}
export interface BatchItem { data: Uint8Array; result: InflateResult; error?: Error; }

const OUTPUT_BUFSIZE = 16384;
const enum Mode { Sniff = 0, Inflater = 1, Raw = 2 }
const CHECK = ["unchecked", "match", "mismatch"] as const;
const MSG = ["", "invalid gzip id", "unknown compression method", "invalid window size", "incorrect header check",
	"need dictionary", "invalid block type", "invalid stored block lengths", "too many length or distance symbols",
	"invalid bit length repeat", "oversubscribed dynamic bit lengths tree", "incomplete dynamic bit lengths tree",
	"oversubscribed literal/length tree", "incomplete literal/length tree", "oversubscribed distance tree",
	"incomplete distance tree", "empty distance tree with lengths", "invalid distance code", "invalid literal/length code"];
const THROWN = ["", "inflate error: bad input", "Custom dictionary is not valid for this data",
	"Custom dictionary required for this data", "inflate error: ", "inflate error: bad input data",
	"reference implementation does not terminate on this input", "data buffer is too small",
	"Unexpected EOF during decompression", "Data integrity check failed", "Data size check failed", "Decompression error"];
const RECORD = 72;   // sizeof(struct sdz_result), include/sdz_codes.h

function u8(source: BufferSource): Uint8Array | undefined {          // src/common.ts:102-114
	if (source instanceof ArrayBuffer) return new Uint8Array(source);
	if (!ArrayBuffer.isView(source)) return undefined;
	return source instanceof Uint8Array ? source : new Uint8Array(source.buffer, source.byteOffset, source.byteLength);
}

export function adler32(source: BufferSource, seed = 1): number {
	const view = u8(source);
	if (!view) throw new TypeError("source must be a BufferSource");
	return native.adler32(view, seed | 0);
}
export function crc32(source: BufferSource, seed = 0): number {
	const view = u8(source);
	if (!view) throw new TypeError("source must be a BufferSource");
	return native.crc32(view, seed | 0);
}
export function mergeBuffers(buffers: Uint8Array[]): Uint8Array {    // src/common.ts:116-126
	const out = new Uint8Array(buffers.reduce((s, b) => s + b.byteLength, 0));
	let off = 0;
	for (const b of buffers) { out.set(b, off); off += b.length; }
	return out;
}

interface Rec { outLen: number; msg: number; thrownAppend: number; thrownInflate: number; result: InflateResult; }
function parseRecord(rec: DataView, input: Uint8Array): Rec {
	const nameOff = rec.getUint32(44, true), nameLen = rec.getUint32(48, true), mtime = rec.getInt32(40, true);
	let fileName = "";
	for (let i = 0; i < nameLen; i++) fileName += String.fromCharCode(input[nameOff + i]);   // Latin-1, src/inflate.ts:387
	return {
		outLen: Number(rec.getBigUint64(8, true)), msg: rec.getUint8(56), thrownAppend: rec.getUint8(57), thrownInflate: rec.getUint8(58),
		result: {
			success: rec.getUint8(63) !== 0, complete: rec.getUint8(60) !== 0,
			checksum: CHECK[rec.getUint8(61)], fileSize: CHECK[rec.getUint8(62)],
			fileName, modDate: mtime === 0 ? undefined : new Date(mtime * 1000),
		},
	};
}
function errorOf(thrown: number, msg: number): Error {
	if (thrown === 12) return new RangeError("options.dictionary cannot be set when options.raw is true");
	return new Error(THROWN[thrown] + (thrown === 4 ? MSG[msg] : ""));
}

function runBatch(views: Uint8Array[], dicts: (Uint8Array | undefined)[], modes: Uint8Array) {
	const n = views.length;
	const sizes = native.inflateSizes(views, dicts, modes);
	const off = new BigUint64Array(n), cap = new BigUint64Array(n);
	let total = 0;
	for (let i = 0; i < n; i++) { off[i] = BigInt(total); const c = (sizes[i] + 15) & ~15; cap[i] = BigInt(c); total += c; }
	const arena = new Uint8Array(total + 64), records = new Uint8Array(n * RECORD);
	native.inflateBatch(views, dicts, modes, arena, off, cap, records);
	return { arena, off, records };
}

/** The one entry point added to the reference API: errors are recorded per stream, never thrown. */
export function inflateBatch(buffers: BufferSource[], options?: { dictionaries?: (BufferSource | undefined)[] }): BatchItem[] {
	const views = buffers.map(b => { const v = u8(b); if (!v) throw new TypeError("data must be an ArrayBuffer or buffer view"); return v; });
	const dicts = views.map((_, i) => { const d = options?.dictionaries?.[i]; return d === undefined ? undefined : u8(d); });
	const { arena, off, records } = runBatch(views, dicts, new Uint8Array(views.length));
	return views.map((v, i) => {
		const r = parseRecord(new DataView(records.buffer, i * RECORD, RECORD), v);
		const data = r.thrownAppend ? new Uint8Array(0) : arena.slice(Number(off[i]), Number(off[i]) + r.outLen);
		return { data, result: r.result, error: r.thrownInflate ? errorOf(r.thrownInflate, r.msg) : undefined };
	});
}

/** inflate(data, dictionary?) - src/sd-inflate.ts:189-228 */
export function inflate(data: BufferSource, dictionary?: BufferSource): Uint8Array {
	const input = u8(data);
	if (!(input instanceof Uint8Array)) throw new TypeError("data must be an ArrayBuffer or buffer view");
	if (input.length < 2) throw new Error("data buffer is too small");
	const [item] = inflateBatch([input], { dictionaries: [dictionary] });
	if (item.error) throw item.error;
	return item.data;
}

/** class Inflater - src/sd-inflate.ts:54-180 over a device session (sdz_inflater_*, include/sdzcuda.h): every append()
 *  decodes only its own chunk; the reference's behaviour at chunk boundaries (Q2 / Q3 / Q4) is reproduced */
export class Inflater {
	private raw: boolean; private dict: Uint8Array | undefined;
	private handle: unknown; private rec = new Uint8Array(RECORD); private used = false;
	constructor(options?: InflaterOptions) {
		const raw = options?.raw;
		if (raw !== undefined && raw !== true && raw !== false) throw new TypeError("options.raw must be undefined or true or false");
		this.raw = raw === true;
		if (options?.dictionary !== undefined) {
			if (this.raw) throw new RangeError("options.dictionary cannot be set when options.raw is true");
			this.dict = u8(options.dictionary);
			if (this.dict === undefined) throw new TypeError("options.dictionary must be undefined or a buffer or a buffer view");
		}
	}
	append(data: BufferSource): Uint8Array[] {
		const chunk = u8(data);
		if (!(chunk instanceof Uint8Array)) throw new TypeError("data must be an ArrayBuffer or buffer view");
		if (chunk.length === 0) return [];
		this.handle ??= native.inflaterNew(this.raw, this.dict);
		const fresh: Uint8Array = native.inflaterAppend(this.handle, chunk, this.rec);
		this.used = true;
		const r = parseRecord(new DataView(this.rec.buffer, 0, RECORD), new Uint8Array(0));
		if (r.thrownAppend) throw errorOf(r.thrownAppend, r.msg);
		const out: Uint8Array[] = [];
		// every append() starts with an empty 16 KiB buffer (src/sd-inflate.ts:101-103): these are the reference's chunk shapes
		for (let o = 0; o < fresh.length; o += OUTPUT_BUFSIZE) out.push(fresh.slice(o, Math.min(fresh.length, o + OUTPUT_BUFSIZE)));
		return out;
	}
	finish(): InflateResult {
		if (!this.used) return { success: false, complete: false, checksum: "unchecked", fileSize: "unchecked", fileName: "", modDate: undefined };
		const name: Uint8Array = native.inflaterFinish(this.handle, this.rec);
		const r = parseRecord(new DataView(this.rec.buffer, 0, RECORD), new Uint8Array(0));
		return { ...r.result, fileName: String.fromCharCode(...name) };
	}
}
