// index.js -- JavaScript shim with the export names of compressjs' main.js for the bzip2 path.
// Argument coercion (streams / buffers / sizes) stays in JavaScript exactly as the reference does it;
// the per-block work goes to the N-API addon (addon.cc -> libb2bz.so -> CUDA).
'use strict';
var native = require('./build/Release/b2bz.node');

var EOF = -1;
function drain(input) {                       // Util.coerceInputStream semantics
  if (input && typeof input.readByte === 'function') {
    var bytes = [], b;
    while ((b = input.readByte()) !== EOF) { bytes.push(b); }
    return Buffer.from(bytes);
  }
  return Buffer.isBuffer(input) ? input : Buffer.from(input);
}
function deliver(output, data) {              // Util.coerceOutputStream + retval semantics
  if (output && typeof output === 'object' && typeof output.writeByte === 'function') {
    for (var i = 0; i < data.length; i++) { output.writeByte(data[i]); }
    return output;
  }
  if (typeof output === 'number') {
    if (output !== data.length) { throw new TypeError('outputsize does not match decoded input'); }
    return new Uint8Array(data);
  }
  if (output) {
    if (output.length !== data.length) { throw new TypeError('outputsize does not match decoded input'); }
    for (var j = 0; j < data.length; j++) { output[j] = data[j]; }
    return output;
  }
  return new Uint8Array(data);
}

var Bzip2 = Object.create(null);
Bzip2.compressFile = function(inStream, outStream, props) {
  var level = (typeof props === 'number') ? props : 9;
  if (level < 1 || level > 9) { throw new Error('Invalid block size multiplier'); }
  return deliver(outStream, native.compressFile(drain(inStream), level));
};
Bzip2.decompressFile = function(input, output, multistream) {
  return deliver(output, native.decompressFile(drain(input), !!multistream));
};
Bzip2.decompressBlock = function(input, pos, output) {
  return deliver(output, native.decompressBlock(drain(input), pos));
};
Bzip2.table = function(input, callback, multistream) {
  native.table(drain(input), !!multistream).forEach(function(r) { callback(r[0], r[1]); });
};

var BWT = Object.create(null);
BWT.bwtransform2 = function(T, U, n) { return native.bwtransform2(T, U, n); };
// the sentinel family (lib/BWT.js:305-363); the reference's scratch arrays A / LF are accepted and ignored
BWT.suffixsort = function(T, SA, n) { return native.suffixsort(T, SA, n); };
BWT.bwtransform = function(T, U, A, n) { return native.bwtransform(T, U, n); };
BWT.unbwtransform = function(T, U, LF, n, pidx) { native.unbwtransform(T, U, n, pidx); };

var BWTC = Object.create(null);   // lib/BWTC.js:10-231
BWTC.MAGIC = 'bwtc';
BWTC.compressFile = function(inStream, outStream, props) {
  var level = (typeof props === 'number' && props >= 1 && props <= 9) ? props : 9;   // :16-19
  return deliver(outStream, native.bwtcCompressFile(drain(inStream), level));
};
BWTC.decompressFile = function(input, output) {
  return deliver(output, native.bwtcDecompressFile(drain(input)));
};

module.exports = Object.freeze({ version: '0.0.1', Bzip2: Bzip2, BWT: BWT, BWTC: BWTC });
